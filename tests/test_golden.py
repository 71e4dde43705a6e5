"""Committed fixtures (tests/golden/ndt2d_golden.npz, made by tests/golden/make_golden.py from the spec oracle).
CPU: the oracle still reproduces them (guards SPEC.md's arithmetic). GPU: the CUDA path reproduces them without
the live oracle in the loop. They are not upstream vectors: the reference mount has no source (parity unpinned)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ndt2d_golden.npz"))


def _scene():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("ov", [0, 1])
def test_oracle_reproduces_golden(ov):
    cur = _scene().compute(ov)
    for k, v in cur.items():
        g = GOLD[f"ov{ov}_{k}"]
        if np.asarray(v).dtype.kind == "f" and k not in ("cells_nz",):
            assert np.allclose(v, g, rtol=1e-12, atol=1e-12), k
        else:
            assert np.array_equal(v, g), k


@pytest.mark.gpu
@pytest.mark.parametrize("ov", [0, 1])
def test_gpu_reproduces_golden(ov):
    import gtsam_ndt_b200 as gn
    from gtsam_ndt_b200 import synth
    map_xy, scans, poses, init = _scene().scene()
    m = gn.NdtMatcher2D([1.0, 0.5], overlap=ov)
    m.set_target(map_xy)
    G = lambda k: GOLD[f"ov{ov}_{k}"]
    cells = m.cells(1)
    nz = np.argwhere(cells[..., 7] != 0)
    assert np.array_equal(nz.astype(np.int32), G("cells_nz_index")) and cells[nz[:, 0], nz[:, 1]].tobytes() == G("cells_nz").tobytes()
    assert np.array_equal(m.cell_index(scans[0], init[0], level=1), G("cell_index"))
    xy, off = synth.pack(scans)
    r = m.align_batch(xy, off, init)
    assert np.array_equal(r["iterations"], G("res_iter")) and np.array_equal(r["status"], G("res_status"))
    assert np.array_equal(r["count"], G("res_count"))
    assert np.abs(r["pose"] - G("res_pose"))[:, :2].max() <= 1e-5 and np.abs(r["pose"] - G("res_pose"))[:, 2].max() <= 1e-6
    assert np.allclose(r["score"], G("res_score"), rtol=1e-6)
    assert np.allclose(r["hessian"], G("res_hessian"), rtol=1e-6, atol=1e-6 * np.abs(G("res_hessian")).max())
    ev, cnt = m.evaluate(scans[3], init[3], level=1)
    assert cnt == G("eval_count")[3] and np.array_equal(ev, G("eval10")[3])
    hyp = (poses[2] + np.stack(np.meshgrid(np.arange(-2, 3) * 0.25, np.arange(-2, 3) * 0.25, np.radians(np.arange(-2, 3) * 2.0),
                                           indexing="ij"), -1).reshape(-1, 3)).astype(np.float32)
    s, bi, _ = m.sweep(scans[2], hyp, k=1, level=0)
    assert bi[0] == G("sweep_best") and np.array_equal(s, G("sweep_scores"))
