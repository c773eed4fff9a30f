// Hardware probe: tcgen05.mma.kind::f16 with an MN-major (time-contiguous) fp16 A operand in SWIZZLE_128B and a K-major fp16
// B operand in SWIZZLE_64B (rows of 32 k = 64 bytes) -- the operand forms a "3xFP16" variant of csrc/ctn_umma.cu would use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I dnn-based_source_separation_b200/csrc -o tools/umma_unit_f16 tools/umma_unit_f16.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ctn_umma_ptx.cuh"

struct Case { int N, K; uint32_t lbo_a, sbo_a; int a_step_groups; const char* name; };

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_probe(const __half* __restrict__ A, const __half* __restrict__ Bm, float* __restrict__ D, Case c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(sm + 64);
  __half* sA = reinterpret_cast<__half*>(sm + 1024);          // 16 KB
  __half* sB = reinterpret_cast<__half*>(sm + 1024 + 16384);  // 16 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 2; i += 128) sA[i] = __float2half(-77.f);
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tptr), 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  // A (128 x K), MN-major SW128 16-bit: atoms of 64 m x 8 k (1024 B): k-row r at r*128 B, 16-byte chunks (8 m) XOR r
  for (int i = tid; i < 128 * c.K; i += 128) {
    const int m = i % 128, k = i / 128;
    const int kg = k >> 3, r = k & 7, atom = m >> 6, chunk = (m & 63) >> 3, e = m & 7;
    const int off_bytes = kg * (int)c.sbo_a + atom * (int)c.lbo_a + r * 128 + ((chunk ^ r) << 4) + e * 2;
    sA[off_bytes / 2] = A[m * c.K + k];
  }
  // B (N x K<=32), K-major SWIZZLE_64B: rows of 64 B, 8-row groups 512 B, 16-byte chunk ^ ((row >> 1) & 3)
  for (int i = tid; i < c.N * c.K; i += 128) {
    const int n = i / c.K, k = i % c.K;
    const int off_bytes = (n >> 3) * 512 + (n & 7) * 64 + ((((k >> 3) ^ ((n >> 1) & 3))) << 4) + (k & 7) * 2;
    sB[off_bytes / 2] = Bm[n * c.K + k];
  }
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    ptx::tc_fence_after();
    // kind::f16: c_format F32 (1<<4), a_format = b_format = 0 (F16), a_major MN (bit 15), b_major K
    const uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int kk = 0; kk < c.K / 16; ++kk) {
      const uint64_t da = ptx::make_smem_desc(ptx::smem_u32(sA) + kk * c.a_step_groups * c.sbo_a, c.lbo_a, c.sbo_a, 2);
      const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(sB) + kk * 32, 16, 512, 4);
      mma_f16(tmem, da, db, idesc, kk ? 1u : 0u);
    }
    ptx::mma_commit(ptx::smem_u32(&bars[0]));
  }
  ptx::mbar_wait(ptx::smem_u32(&bars[0]), 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < c.N; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * c.N + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

static bool run_case(const Case& c) {
  const int M = 128;
  std::vector<__half> A(M * c.K), B(c.N * c.K);
  std::vector<float> Dref(M * c.N, 0.f), D(M * c.N, -1.f);
  for (int m = 0; m < M; ++m) for (int k = 0; k < c.K; ++k) A[m * c.K + k] = __float2half(float((m + 3 * k) % 13 - 6));
  for (int n = 0; n < c.N; ++n) for (int k = 0; k < c.K; ++k) B[n * c.K + k] = __float2half(float((n * 5 + k) % 7 - 3));
  for (int m = 0; m < M; ++m) for (int n = 0; n < c.N; ++n) {
    float s = 0.f;
    for (int k = 0; k < c.K; ++k) s += __half2float(A[m * c.K + k]) * __half2float(B[n * c.K + k]);
    Dref[m * c.N + n] = s;
  }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  const size_t smem = 1024 + 1024 + 32768;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_probe<<<1, 128, smem>>>(dA, dB, dD, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", c.name, cudaGetErrorString(e)); exit(2); }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - Dref[i]); if (d > 0 || d != d) ++bad; if (d > maxerr || d != d) maxerr = d; }
  printf("[%-40s] N=%d K=%d lbo_a=%u sbo_a=%u step=%d : mismatches %d/%zu maxerr %.1f\n", c.name, c.N, c.K, c.lbo_a, c.sbo_a,
         c.a_step_groups, bad, D.size(), maxerr);
  if (bad) {
    printf("    D[0][0..7] got:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", D[j]); printf("\n    D[0][0..7] ref:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", Dref[j]);
    printf("\n    D[70][0..7] got:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", D[70 * c.N + j]); printf("\n    D[70][0..7] ref:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", Dref[70 * c.N + j]);
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad == 0;
}

int main() {
  std::vector<Case> cases = {
      {32, 16, 1024, 2048, 2, "f16 MN A: lbo=1024 sbo=2048, K=16"},
      {32, 32, 1024, 2048, 2, "f16 MN A: lbo=1024 sbo=2048, K=32"},
      {256, 32, 1024, 2048, 2, "f16 MN A: lbo=1024 sbo=2048, N=256 K=32"},
      {32, 16, 2048, 1024, 2, "f16 MN A: fields swapped, K=16"},
      {32, 32, 4096, 1024, 2, "f16 MN A: k-groups adjacent (lbo=4096 sbo=1024)"},
  };
  int ok = 0;
  for (const Case& c : cases) ok += run_case(c);
  printf("passed %d / %zu\n", ok, cases.size());
  return 0;
}
