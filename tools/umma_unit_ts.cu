// Hardware probe: tcgen05.mma.kind::f16 with the A operand in TENSOR MEMORY (the "ts" form) and a K-major fp16 B operand in
// SWIZZLE_64B shared memory.  Checks the TMEM layout of a 16-bit A (lane = row m, 32-bit column j = {A[m][2j], A[m][2j+1]}),
// written with tcgen05.st by the warp that owns the lanes, and times a chain of such MMAs (cycles per instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I dnn-based_source_separation_b200/csrc -o tools/umma_unit_ts tools/umma_unit_ts.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ctn_umma_ptx.cuh"

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A: (128 x K) row-major fp16 in global; B: (N x K) row-major; D: (128 x N) fp32.  reps > 0: timing mode (result garbage)
__global__ void __launch_bounds__(128) k_probe(const __half* __restrict__ A, const __half* __restrict__ Bm, float* __restrict__ D, int N,
                                               int K, int reps, long long* cyc, int ss_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(sm + 64);
  __half* sB = reinterpret_cast<__half*>(sm + 1024);                 // K/32 slabs of N x 64 B
  __half* sA = reinterpret_cast<__half*>(sm + 1024 + 98304);         // ss_mode: K/32 slabs of 128 x 64 B (K-major SW64)
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tptr), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t tmem_acc = tmem, tmem_a = tmem + 256;               // accumulator cols [0,256), A cols [256, 256 + K/2)
  // B, K-major SWIZZLE_64B slabs of 32 k
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K, slab = k >> 5, kk = k & 31;
    const int off_bytes = slab * (N * 64) + (n >> 3) * 512 + (n & 7) * 64 + ((((kk >> 3) ^ ((n >> 1) & 3))) << 4) + (kk & 7) * 2;
    sB[off_bytes / 2] = Bm[n * K + k];
  }
  for (int i = tid; i < 128 * K; i += 128) {
    const int n = i / K, k = i % K, slab = k >> 5, kk = k & 31;
    const int off_bytes = slab * (128 * 64) + (n >> 3) * 512 + (n & 7) * 64 + ((((kk >> 3) ^ ((n >> 1) & 3))) << 4) + (kk & 7) * 2;
    sA[off_bytes / 2] = A[n * K + k];
  }
  // A into TMEM: thread tid owns lane tid; 8 columns (16 k) per store
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      const __half lo = A[tid * K + 2 * (c0 + j)], hi = A[tid * K + 2 * (c0 + j) + 1];
      v[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    tmem_st8(tmem_a + ((uint32_t)(warp * 32) << 16) + c0, v);
  }
  tmem_st_wait();
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0 && ptx::elect_one()) {
    ptx::tc_fence_after();
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A, B K-major
    const long long t0 = clock64();
    const int R = reps > 0 ? reps : 1;
    const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 512, 4);
    const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(sA), 16, 512, 4);
    const uint32_t bslab = (uint32_t)(N * 64) >> 4, aslab = (128 * 64) >> 4;
    const int nacc = 1 << (ss_mode >> 1);        // bits 1..: round-robin over 1 / 2 / 4 accumulators PER INSTRUCTION (timing only)
    const uint32_t accw = 256u / nacc;
    int q = 0;
    for (int r = 0; r < R; ++r) {
#pragma unroll 1
      for (int sl = 0; sl < K / 32; ++sl) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t db = db0 + (uint64_t)(sl * bslab + h * 2);
          const uint32_t dacc = tmem_acc + (reps > 0 ? (uint32_t)(q & (nacc - 1)) * accw : 0u);
          const uint32_t acc = (reps > 0) ? (q >= nacc ? 1u : 0u) : ((sl || h) ? 1u : 0u);
          ++q;
          if (ss_mode & 1) ptx::mma_f16(dacc, da0 + (uint64_t)(sl * aslab + h * 2), db, idesc, acc);
          else mma_f16_ts(dacc, tmem_a + (sl * 2 + h) * 8, db, idesc, acc);
        }
      }
    }
    const long long t1 = clock64();
    ptx::mma_commit(ptx::smem_u32(&bars[0]));
    ptx::mbar_wait(ptx::smem_u32(&bars[0]), 0);
    if (cyc) { cyc[0] = clock64() - t0; cyc[1] = t1 - t0; }
  }
  ptx::mbar_wait(ptx::smem_u32(&bars[0]), 0);
  ptx::tc_fence_after();
  if (reps == 0)
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld16(tmem_acc + ((uint32_t)(warp * 32) << 16) + c0, v);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

static bool run_case(int N, int K, int ss) {
  const int M = 128;
  std::vector<__half> A(M * K), B(N * K);
  std::vector<float> Dref(M * N, 0.f), D(M * N, -1.f);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = __float2half(float((m + 3 * k) % 13 - 6));
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = __float2half(float((n * 5 + k) % 7 - 3));
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += __half2float(A[m * K + k]) * __half2float(B[n * K + k]);
    Dref[m * N + n] = s;
  }
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  const size_t smem = 1024 + 1024 + 98304 + 49152;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_probe<<<1, 128, smem>>>(dA, dB, dD, N, K, 0, nullptr, ss);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[N=%d K=%d ss=%d] CUDA error: %s\n", N, K, ss, cudaGetErrorString(e)); exit(2); }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - Dref[i]); if (d > 0 || d != d) ++bad; if (d > maxerr || d != d) maxerr = d; }
  printf("[%s%s N=%3d K=%3d] mismatches %d/%zu maxerr %.1f", (ss & 1) ? "SS" : "TS", (ss >> 1) == 2 ? "/4acc" : ((ss >> 1) == 1 ? "/2acc" : ""), N, K, bad, D.size(), maxerr);
  if (bad) {
    printf("\n    D[0][0..7] got:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", D[j]); printf("\n    D[0][0..7] ref:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", Dref[j]);
    printf("\n    D[70][0..7] got:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", D[70 * N + j]); printf("\n    D[70][0..7] ref:"); for (int j = 0; j < 8; ++j) printf(" %6.1f", Dref[70 * N + j]);
  }
  // timing: 64 repetitions of the K/16-instruction chain
  for (int reps : {1, 2, 4, 64}) {
    long long c[2] = {0, 0};
    for (int it = 0; it < 2; ++it) {
      k_probe<<<1, 128, smem>>>(dA, dB, dD, N, K, reps, dC, ss);
      cudaDeviceSynchronize();
      cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
    }
    printf("   | %d MMAs: issue returned after %lld cycles, complete after %lld (%.1f cycles/MMA)", reps * K / 16, c[1], c[0], (double)c[0] / (reps * K / 16));
  }
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad == 0;
}

int main() {
  int ok = 0, n = 0;
  const int Ns[] = {16, 64, 128, 256};
  for (int ss = 0; ss < 2; ++ss)
    for (int N : {64, 128, 256})
      for (int K : {192}) { ok += run_case(N, K, ss); ++n; }
  printf("passed %d / %d\n", ok, n);
  return 0;
}
