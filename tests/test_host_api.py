"""CPU-side tests: the C-ABI library loads and exports every symbol include/ctn_b200.h declares, host geometry /
workspace logic, error mapping, module tree + state_dict parity with the reference's key list, no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

import convtasnet_oracle as O
from ctn_b200 import _native as N
from ctn_b200.models.conv_tasnet import ConvTasNet, Separator
from ctn_b200.models.tdcn import TimeDilatedConvNet
from ctn_b200.models.filterbank import Encoder, Decoder
from ctn_b200.modules.norm import GlobalLayerNorm, CumulativeLayerNorm1d
from ctn_b200.criterion.sdr import NegSISDR, SISDR
from ctn_b200.criterion.pit import PIT1d, pit
from ctn_b200.utils.tasnet import choose_layer_norm
from ctn_b200.utils.filterbank import choose_filterbank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "ctn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ctn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(N.lib, name), f"{name} declared in ctn_b200.h but not exported by libctn_b200.so"
    assert declared == set(N.EXPORTED)
    assert N.ctn_version() == 100
    assert b"envelope" in N.ctn_strerror(N.CTN_EUNSUPPORTED)


@pytest.mark.parametrize("T,L,S", [(32000, 16, 8), (203, 4, 2), (2500, 16, 8), (128000, 16, 8), (31, 16, 8), (17, 16, 8),
                                   (16, 16, 8), (100, 2, 1), (101, 20, 10)])
def test_frames_matches_reference_rule(T, L, S):
    # src/models/conv_tasnet.py:145-147
    padding = (S - (T - L) % S) % S
    pl = padding // 2
    pr = padding - pl
    frames = (T + padding - L) // S + 1
    assert N.frames_of(T, L, S) == (frames, pl, pr)
    assert N.ctn_pitch(frames) % 128 == 0 and 0 <= N.ctn_pitch(frames) - frames < 128


def test_frames_paper_configs():
    assert N.frames_of(32000, 16, 8)[0] == 3999      # cfg2
    assert N.frames_of(128000, 16, 8)[0] == 15999    # cfg5
    with pytest.raises(ValueError):
        N.frames_of(10, 16, 7)                        # kernel % stride != 0


def _cfg(**kw):
    c = N.Config()
    base = dict(n_basis=512, kernel_size=16, stride=8, bottleneck=128, hidden=512, skip=128, sep_kernel=3, num_blocks=3,
                num_layers=8, n_sources=2, causal=0, enc_relu=0, mask_softmax=0, math=0, eps=1e-12, eps_tcn=1e-12)
    base.update(kw)
    for k, v in base.items():
        setattr(c, k, v)
    return c


def test_workspace_bytes_and_envelope():
    need = C.c_size_t(0)
    c = _cfg()
    assert N.ctn_workspace_bytes(C.byref(c), 32, 32000, C.byref(need)) == 0
    b32 = need.value
    assert N.ctn_workspace_bytes(C.byref(c), 16, 32000, C.byref(need)) == 0
    assert 0 < need.value < b32 < 8 << 30
    # per-sample activations at pitch 4096: w(512)+what(1024)+x(128)+skip(128)+h(512)+u(512) rows + 24 blocks x r(256)
    rows = 512 + 1024 + 128 + 128 + 512 + 512 + 24 * 256
    assert b32 >= 32 * rows * 4096 * 4
    cb = _cfg(mask_softmax=1)   # softmax masks: forward built (logits + one normalising pass), training path not
    assert N.ctn_workspace_bytes(C.byref(cb), 1, 32000, C.byref(need)) == 0 and need.value > 0
    assert N.ctn_train_workspace_bytes(C.byref(cb), 1, 32000, C.byref(need)) == N.CTN_EUNSUPPORTED
    cc = _cfg(causal=1)   # cLN models: forward built (un-fused pipeline), training path not
    assert N.ctn_workspace_bytes(C.byref(cc), 1, 32000, C.byref(need)) == 0 and need.value > 0
    assert N.ctn_train_workspace_bytes(C.byref(cc), 1, 32000, C.byref(need)) == N.CTN_EUNSUPPORTED
    assert N.ctn_train_workspace_bytes(C.byref(c), 32, 32000, C.byref(need)) == 0 and 10 << 30 < need.value < 40 << 30
    cb = _cfg(kernel_size=16, stride=7)
    assert N.ctn_workspace_bytes(C.byref(cb), 1, 32000, C.byref(need)) == N.CTN_EINVAL
    with pytest.raises(NotImplementedError):
        N.check(N.CTN_EUNSUPPORTED, "x")
    with pytest.raises(ValueError):
        N.check(N.CTN_EINVAL, "x")
    with pytest.raises(RuntimeError):
        N.check(N.CTN_EWORKSPACE, "x")
    with pytest.raises(RuntimeError):
        N.check(700, "cuda error code")


def test_null_pointer_rejected_without_gpu():
    # argument validation happens before any CUDA call
    assert N.ctn_encoder_fwd(None, None, None, 1, 100, 0, 0, 8, 16, 8, 0, 128, None, None) == N.CTN_EINVAL
    assert N.ctn_sisdr_pit_fwd(None, None, 1, 2, 100, 1e-12, None, None, None, None, None, None) == N.CTN_EINVAL


def _paper(n_sources=2, causal=False, **kw):
    return ConvTasNet(512, 16, enc_basis='trainable', dec_basis='trainable', enc_nonlinear=None, sep_hidden_channels=512,
                      sep_bottleneck_channels=128, sep_skip_channels=128, sep_kernel_size=3, sep_num_blocks=3,
                      sep_num_layers=8, causal=causal, n_sources=n_sources, **kw)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("n_sources", [2, 3])
def test_state_dict_matches_reference_key_list(causal, n_sources):
    m = _paper(n_sources=n_sources, causal=causal)
    cfg = O.OracleConfig(n_sources=n_sources, causal=causal)
    spec = [(k, tuple(s)) for k, s in O.state_dict_spec(cfg)]   # verified against the reference by make_golden.py
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == spec
    m.load_state_dict(O.synth_state_dict(cfg, seed=3), strict=True)
    if not causal and n_sources == 2:
        assert m.num_parameters == 4984881 and len(spec) == 343


def test_get_config_build_model_roundtrip(tmp_path):
    m = _paper(n_sources=3)
    cfg = m.get_config()
    assert cfg['stride'] == 8 and cfg['n_sources'] == 3 and cfg['enc_nonlinear'] is None and cfg['causal'] is False
    package = dict(cfg)
    package['state_dict'] = m.state_dict()
    package['n_bases'] = package.pop('n_basis')      # legacy key aliases tolerated (conv_tasnet.py:204-206)
    path = tmp_path / "ckpt.pth"
    torch.save(package, path)
    m2 = ConvTasNet.build_model(str(path), load_state_dict=True)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_reference_trainer_checkpoint_loads(golden_dir):
    """A checkpoint written by the UNMODIFIED reference the way its trainer does (driver.py:208-226; tests/golden/make_golden.py
    checkpoint_case) rebuilds through build_model(load_state_dict=True): same config, same 71 tensors; the resumed optimizer state of
    the reference (`optim_dict`) loads into torch.optim.Adam over our parameters (same parameter order)."""
    path = os.path.join(golden_dir, "ref_ckpt_tiny_gln.pth")
    m = ConvTasNet.build_model(path, load_state_dict=True)
    pkg = torch.load(path, map_location="cpu", weights_only=False)
    assert m.get_config() == {k: pkg[k] for k in m.get_config()}
    sd = O.synth_state_dict(O.OracleConfig(n_basis=16, kernel_size=4, sep_hidden_channels=16, sep_bottleneck_channels=8, sep_skip_channels=8,
                                           sep_num_blocks=2, sep_num_layers=3, causal=False), seed=11)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.load_state_dict(pkg['optim_dict'])          # trainer resume path (driver.py:51-68)
    assert pkg['epoch'] == 1
    with pytest.raises(FileNotFoundError):
        ConvTasNet.build_from_pretrained(root=str(golden_dir), task='wsj0-mix', sample_rate=8000, n_sources=2)
    with pytest.raises(KeyError):
        ConvTasNet.build_from_pretrained(task='no-such-task')


def test_constructor_envelope_errors():
    with pytest.raises(AssertionError):
        ConvTasNet(64, 16, stride=7, enc_basis='trainable', dec_basis='trainable', enc_nonlinear=None)
    with pytest.raises(NotImplementedError):
        ConvTasNet(64, 16, enc_basis='Fourier', dec_basis='Fourier', enc_nonlinear=None, window_fn='hann',
                   enc_onesided=True, enc_return_complex=True)
    with pytest.raises(NotImplementedError):
        ConvTasNet(64, 16, enc_basis='trainable', dec_basis='pinv')
    assert _paper(mask_nonlinear='softmax').separator.mask_softmax is True
    with pytest.raises(ValueError):
        _paper(mask_nonlinear='tanh')
    with pytest.raises(NotImplementedError):
        _paper(separable=False)
    with pytest.raises(NotImplementedError):
        _paper(dilated=False)
    with pytest.raises(NotImplementedError):
        _paper(sep_nonlinear=None)
    with pytest.raises(ValueError):
        choose_layer_norm('gLN', 8, causal=True)
    with pytest.raises(NotImplementedError):
        choose_layer_norm('foo', 8)
    assert Encoder(2, 8).conv1d.weight.shape == (8, 2, 16)      # multichannel filter banks are built (forward only)
    with pytest.raises(NotImplementedError):
        Encoder(65, 8)
    m2 = ConvTasNet(64, 16, enc_basis='trainable', dec_basis='trainable', enc_nonlinear=None, in_channels=2)
    assert m2.native_config().in_channels == 2 and m2.get_config()['in_channels'] == 2
    assert m2.decoder.conv_transpose1d.weight.shape == (64, 2, 16)
    assert isinstance(choose_layer_norm('cLN', 8, causal=True), CumulativeLayerNorm1d)
    assert isinstance(choose_layer_norm('gLN', 8), GlobalLayerNorm)
    enc, dec = choose_filterbank(32, 16, stride=8, enc_basis='trainable', dec_basis='trainable', enc_nonlinear='relu')
    assert enc.nonlinear is True and dec.get_basis().shape == (32, 1, 16)


def test_no_cpu_fallback():
    m = ConvTasNet(16, 4, enc_basis='trainable', dec_basis='trainable', enc_nonlinear=None, sep_hidden_channels=16,
                   sep_bottleneck_channels=8, sep_skip_channels=8, sep_num_blocks=1, sep_num_layers=2, causal=False)
    x = torch.randn(1, 1, 64)
    with torch.no_grad():
        for fn in (lambda: m(x), lambda: m.encoder(x), lambda: m.decoder(torch.randn(1, 16, 10)),
                   lambda: GlobalLayerNorm(4)(torch.randn(1, 4, 9)), lambda: CumulativeLayerNorm1d(4)(torch.randn(1, 4, 9)),
                   lambda: NegSISDR()(torch.randn(2, 2, 64), torch.randn(2, 2, 64)),
                   lambda: m.separator(torch.randn(1, 16, 20)), lambda: m.separator.tdcn(torch.randn(1, 8, 20))):
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                fn()
    with pytest.raises(ValueError):
        with torch.no_grad():
            m(torch.randn(4, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):   # training path: native too, never an eager fallback
        m(x)
    with pytest.raises(NotImplementedError):   # pieces without backward kernels stay loud under autograd
        m.separator(torch.randn(1, 16, 20))


def test_generic_pit_loop_matches_oracle_on_cpu():
    # non-fused criterion path (host logic, reference semantics pit.py:9-44)
    class L1(torch.nn.Module):
        maximize = False

        def forward(self, input, target, batch_mean=True):
            l = (input - target).abs().mean(dim=(1, 2))
            return l.mean() if batch_mean else l
    g = torch.Generator().manual_seed(0)
    t = torch.randn(3, 3, 50, generator=g)
    e = t[:, [2, 0, 1]] + 0.01 * torch.randn(3, 3, 50, generator=g)
    loss, pattern = PIT1d(L1(), 3)(e, t, batch_mean=False)
    assert pattern.tolist() == [[2, 0, 1]] * 3 and loss.shape == (3,)
    assert PIT1d(NegSISDR(), 3).patterns.tolist() == [list(p) for p in __import__("itertools").permutations(range(3))]
    assert NegSISDR().maximize is False and SISDR().maximize is True


def test_dropin_shims_resolve():
    pkg = os.path.join(ROOT, "dnn-based_source_separation_b200")
    code = ("import warnings; warnings.simplefilter('ignore');"
            "from models.conv_tasnet import ConvTasNet; from models.tdcn import TimeDilatedConvNet;"
            "from models.tcn import TemporalConvNet; from models.filterbank import Encoder, Decoder;"
            "from modules.norm import GlobalLayerNorm, CumulativeLayerNorm1d; from norm import GlobalLayerNorm as G2;"
            "from criterion.sdr import NegSISDR, sisdr; from criterion.pit import PIT1d, pit;"
            "from utils.tasnet import choose_layer_norm; import ctn_b200.models.conv_tasnet as m;"
            "assert ConvTasNet is m.ConvTasNet and issubclass(TemporalConvNet, TimeDilatedConvNet) and G2 is GlobalLayerNorm;"
            "print('ok')")
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=pkg), capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_training_param_slots_cover_every_parameter():
    """The autograd node passes the parameters to C in a fixed slot order (ctn_params_t / ctn_block_params_t); every
    nn.Parameter of the model must appear exactly once, and nothing else."""
    from ctn_b200.models._train import param_list, TOP_FIELDS
    m = _paper(n_sources=3)
    slots = param_list(m)
    tensors = [t for _, t in slots if t is not None]
    assert len(tensors) == len(list(m.parameters())) == 343
    assert {id(t) for t in tensors} == {id(p) for p in m.parameters()}
    assert [s for s, _ in slots[:len(TOP_FIELDS)]] == list(TOP_FIELDS)
    # the last residual block has no output head (tdcn.py:58-61): its two slots are None, all others are tensors
    none_slots = [s for s, t in slots if t is None]
    assert none_slots == [(23, "out_w"), (23, "out_b")]
    assert N.MATH_NAMES["f16x3"] == 3 and N.MATH_NAMES["tf32x3"] == 1
