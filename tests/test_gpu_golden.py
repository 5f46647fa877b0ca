"""The CUDA path against the committed golden fixtures recorded from the UNMODIFIED reference
(tests/golden/*.npz: compress.py + decompress.py + prednet.py, made by tests/golden/make_golden.py;
tests/golden/prednet/*.npz: prednet.py alone, made by make_prednet_golden.py).  All calls go through the C ABI.

  * codec kernels GIVEN the reference's recorded predictions: key plane, int16 stream and decoded frames must be
    byte-identical to the reference's files (bit-exact for integer work);
  * the tcgen05 predictor against the reference's own PredNet class: |pred - reference| <= 6e-3 on [0, 1]
    (fp16 operands, fp32 accumulation; the fp32 direct kernels: <= 2e-5)."""
import glob
import os

import numpy as np
import pytest

from test_oracle_golden import GOLD, load
from test_prednet_golden import GOLD as PGOLD, load as pload

pytestmark = pytest.mark.gpu

TOL_TC = 6e-3
TOL_DIRECT = 2e-5


def _schedule(g):
    """pred_slot / apply_eb of the reference's schedule: slot f+1 holds the recorded prediction of frame f."""
    nt, p = g["nt"], g["p"]
    pred_slot = np.arange(1, nt + 1, dtype=np.int32)
    apply_eb = np.ones(nt, np.uint8)
    for wi, (first, n) in enumerate(g["windows"]):
        pred_slot[int(first)] = -1
        apply_eb[int(first)] = 0
        if p != 0 and wi == 0:
            apply_eb[int(first):int(first) + int(n)] = 0          # compress.py:315: no error_bound on the warm-up window
    return pred_slot, apply_eb


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_codec_kernels_reproduce_reference_files(cuda_lib, path):
    import torch
    from tezip_b200 import codec, ops
    g = load(path)
    dev = torch.device("cuda", 0)
    nt, H, W = g["nt"], g["H"], g["W"]
    pool = torch.from_numpy(np.concatenate([g["preds"][:1], g["preds"]], axis=0)).to(dev)
    pred_slot, apply_eb = _schedule(g)
    fr = torch.from_numpy(g["frames"]).to(dev)
    enc = codec.encode_with_pool(fr, pool, pred_slot, apply_eb, [int(k) for k in g["keys"]], g["p"], g["mode"],
                                 g["bound"], g["entropy"], keep_x=True)
    assert np.array_equal(enc.key_plane.cpu().numpy().ravel(), g["ref_key_plane"])
    assert np.array_equal(enc.x.cpu().numpy().ravel(), g["x"])
    assert np.array_equal(enc.payload(), g["ref_payload"])                    # entropy.dat before zstd, byte for byte
    # decoder kernels on the REFERENCE's files
    body, table, shape, pp = codec.parse_payload(g["ref_payload"])
    assert shape == (1, nt, H, W, 3) and pp == g["p"]
    kp = torch.from_numpy(g["ref_key_plane"].reshape(nt, H, W, 3).copy()).to(dev)
    nz = ops.frames_nonzero(kp).cpu().numpy()
    assert [int(i) for i in np.nonzero(nz)[0]] == [int(k) for k in g["keys"]]    # decompress.py:123-127
    lut = torch.from_numpy(ops.decode_lut(table)).to(dev) if table is not None else None
    out = ops.reconstruct(torch.from_numpy(np.ascontiguousarray(body)).to(dev), (nt, H, W, 3), pool.shape[1],
                          pool.shape[2], len(table) if table is not None else -1, lut, pool,
                          torch.from_numpy(pred_slot).to(dev), kp)
    assert np.array_equal(out.cpu().numpy(), g["ref_decoded"])


@pytest.mark.parametrize("path", GOLD, ids=lambda p: os.path.basename(p)[:-4])
def test_whole_gpu_path_against_reference_run(cuda_lib, path):
    """Everything on the GPU (tcgen05 predictions, not the recorded ones): same key placement as the reference, the
    decoded frames within the bound of the originals, lossless cases exact, and a stream of the same size class
    (quantised residuals equal to the reference's except where trunc(pred*255) sits on an integer edge)."""
    import torch
    from tezip_b200 import codec, synth
    from tezip_b200.prednet import PredNet
    g = load(path)
    if g["threshold"] is not None:
        pytest.skip("DWP key placement under fp16 predictions is covered by test_gpu_e2e / test_gpu_dwp")
    dev = torch.device("cuda", 0)
    Hp, Wp = g["preds"].shape[1:3]
    ws = synth.make_weights(g["stack"], bias=g["bias"], seed=7)
    net = PredNet(g["stack"], g["stack"], weights=ws, input_hw=(Hp, Wp), max_batch=4)
    enc = codec.encode_frames(torch.from_numpy(g["frames"]).to(dev), net, g["p"], g["window"], None, g["mode"],
                              g["bound"], g["entropy"], keep_x=True, keep_pool=True)
    assert enc.keys == [int(k) for k in g["keys"]]
    assert np.array_equal(enc.key_plane.cpu().numpy().ravel(), g["ref_key_plane"])
    slot = enc.pred_slot
    used = slot >= 1
    got = enc.pool[torch.from_numpy(slot[used].astype(np.int64)).to(dev)].cpu().numpy()
    assert np.abs(got - g["preds"][used]).max() <= TOL_TC
    if codec.is_lossless(g["mode"], g["bound"]):      # (a lossy segment value moves as a whole when one residual flips)
        assert np.mean(enc.x.cpu().numpy().ravel() == g["x"]) >= 0.97
    out, _plan = codec.decode_arrays(enc.key_plane, enc.body, enc.table, enc.shape, g["p"], net)
    out = out.cpu().numpy()
    if codec.is_lossless(g["mode"], g["bound"]):
        assert np.array_equal(out, g["frames"])
    elif g["mode"] == "abs":
        assert np.abs(out.astype(int) - g["frames"].astype(int)).max() <= int(np.floor(g["bound"][0])) + 1
    net.close()


@pytest.mark.parametrize("path", PGOLD, ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("direct", [False, True], ids=["tcgen05", "fp32direct"])
def test_predictor_against_reference_prednet(cuda_lib, path, direct):
    import torch
    from tezip_b200.prednet import PredNet
    g = pload(path)
    tol = TOL_DIRECT if direct else TOL_TC
    net = PredNet(g["stack"], g["stack"], weights=g["weights"], input_hw=(g["Hp"], g["Wp"]), max_batch=g["B"],
                  fp32_direct=direct)
    dev = torch.device("cuda", 0)
    assert np.abs(net.p0().cpu().numpy() - g["p0"]).max() <= TOL_DIRECT           # create-time fp32 kernels
    n1 = net.next(torch.from_numpy(g["frames"]).to(dev))
    e1 = float(np.abs(n1.cpu().numpy() - g["next1"]).max())
    assert e1 <= tol, e1
    n2 = net.next_chained(torch.empty_like(n1))
    e2 = float(np.abs(n2.cpu().numpy() - g["next2"]).max())
    assert e2 <= 2 * tol, e2
    # from the reference's own fed-back input
    n2r = net.next(torch.from_numpy(g["next1"]).to(dev))
    assert float(np.abs(n2r.cpu().numpy() - g["next2"]).max()) <= tol
    net.close()
