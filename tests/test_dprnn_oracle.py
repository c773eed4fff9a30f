"""CPU: the DPRNN-TasNet oracle (oracle/dprnn_oracle.py) against fixtures minted from the unmodified reference
(tests/golden/make_golden.py: dprnn_cases) -- Segment1d / OverlapAdd1d, a tiny model, and the cfg4 hyper-parameters."""
import os

import pytest
import torch

import convtasnet_oracle as O
import dprnn_oracle as DO


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def test_segment_overlap_add(golden_dir):
    rec = _load(golden_dir, "dprnn_modules")
    assert len(rec) == 3
    for key, r in rec.items():
        _, B, Fc, T, K, P = key.split("_")
        seg = DO.segment1d(r["x"], int(K), int(P))
        assert torch.equal(seg, r["seg"]), key
        torch.testing.assert_close(DO.overlap_add1d(seg, int(K), int(P)), r["ola"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", ["dprnn_tiny", "dprnn_cfg4_short"])
def test_dprnn_tasnet_cases(golden_dir, name):
    rec = _load(golden_dir, name)
    cfg = DO.DPRNNConfig(**rec["cfg"])
    sd = DO.synth_state_dict(cfg, seed=rec["wseed"])
    mixture, sources = O.synth_batch(rec["batch"], cfg.n_sources, rec["T"], seed=rec["xseed"])
    with torch.no_grad():
        out, latent = DO.dprnn_tasnet_fwd(mixture, sd, cfg)
        loss, perm = O.pit_neg_sisdr(out, sources)
    assert torch.equal(perm, rec["perm"])
    torch.testing.assert_close(loss, rec["loss"], rtol=0, atol=1e-4)
    so = rec.get("out_stride")
    if so is None:
        torch.testing.assert_close(out, rec["out"], rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(latent, rec["latent"], rtol=1e-5, atol=2e-6)
    else:
        torch.testing.assert_close(out[..., ::so], rec["out"], rtol=1e-5, atol=5e-6)
        a, b = rec["latent_stride"]
        torch.testing.assert_close(latent[:, :, ::a, ::b], rec["latent"], rtol=1e-5, atol=5e-6)
        assert abs(float(out.double().sum()) - rec["out_sum"]) < 1e-3 * max(1.0, rec["out_sumsq"] ** 0.5)
