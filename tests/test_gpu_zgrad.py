"""GPU parity for SURVEY §8 row f1: gradients of the hot-path operators with respect to the inducing points Z
(lip_zgrad) against the float64 oracle, which differentiates a literal torch restatement of src/ggn.py with
torch.func.grad (what jax.value_and_grad does to the reference, train_inducing.py:195-232), and the native
cross-Gram (lip_gram_cross, build_WTWz of ggn.py:233-272).  Tolerance: 1e-5 relative, as for the products."""
import numpy as np
import pytest
import torch

from helpers import make_pair, rel_err
from oracle import lip_oracle as O
from test_gpu_parity import CONFIGS, _setup, cu

pytestmark = pytest.mark.gpu

TOL = 1e-5
NAMES = ["C1_toy_sine", "C2_xor", "C3a_subset89", "mlp_ragged", "reg_logvar", "linear"]


def _probes(D, B, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, D)).astype(np.float32), rng.standard_normal((B, D)).astype(np.float32)


@pytest.mark.parametrize("name", NAMES)
def test_ggn_vp_zgrad_matches_oracle(name):
    from lip_b200 import ggn, lla
    ost, lst, Z, mt, N = _setup(name)
    D = ost.flat()[0].size
    U, V = _probes(D, 3, 11)
    ref = O.ggn_vp_zgrad(ost, Z, mt, U, V, full_set_size=N, per_probe=True)
    vp = ggn.compute_ggn_vp(lst, cu(Z), mt, full_set_size=N)
    got = vp.zgrad(cu(U), cu(V)).cpu().numpy()
    assert got.shape == Z.shape
    assert rel_err(got, ref.sum(0)) < TOL
    per = vp.zgrad(cu(U), cu(V), per_probe=True).cpu().numpy()
    assert per.shape == (3,) + Z.shape and rel_err(per, ref) < TOL
    # curvature_vp = ggn_vp + alpha v: the alpha term has no Z dependence
    cvp = lla.compute_curvature_approx(lst, cu(Z), mt, 0.3, full_set_size=N)
    assert rel_err(cvp.zgrad(cu(U), cu(V)).cpu().numpy(), ref.sum(0)) < TOL
    # gradient of the quadratic form v^T GGN(Z) v (Hutchinson-style estimators): ubar = v
    q = vp.zgrad(cu(V), cu(V)).cpu().numpy()
    assert rel_err(q, O.ggn_vp_zgrad(ost, Z, mt, V, V, full_set_size=N)) < TOL


@pytest.mark.parametrize("name", NAMES)
def test_W_WT_zgrad_match_oracle(name):
    from lip_b200 import ggn
    ost, lst, Z, mt, N = _setup(name)
    D = ost.flat()[0].size
    M, K = Z.shape[0], CONFIGS[name][2]
    U, V = _probes(D, 3, 12)
    Y = np.random.default_rng(13).standard_normal((3, M, K)).astype(np.float32)
    Wz_ref, WTz_ref = O.W_vps_zgrad(ost, Z, mt, full_set_size=N)
    Wfun, WTfun = ggn.compute_W_vps(lst, cu(Z), mt, full_set_size=N)
    assert rel_err(Wfun.zgrad(cu(U), cu(Y)).cpu().numpy(), Wz_ref(U, Y)) < TOL
    assert rel_err(WTfun.zgrad(cu(Y), cu(V)).cpu().numpy(), WTz_ref(Y, V)) < TOL


TC_ZCONFIGS = {
    # name: (hidden, n_out, in_dim, M, N, activation model)
    "tc_ragged": ([200, 72, 136], 7, 100, 150, 900),
    "tc_m50": ([256, 128, 64], 10, 784, 50, 60000),
    "tc_simt_bottom": ([128, 64], 5, 45, 70, 500),      # first layer (in = 45) and head on SIMT, a tensor-core layer between them
    "tc_top": ([128], 64, 96, 80, 800),                 # the logit layer itself on the tensor cores (K = 64 outputs)
}


@pytest.mark.parametrize("name", list(TC_ZCONFIGS))
def test_zgrad_tensor_core_path_matches_oracle_and_simt(name):
    """lip_zgrad on the tcgen05 path (layers with in, out >= 64: JVP sweep forward, delta-backprop + dual-K GEMMs with the
    mask + add epilogue in reverse) against the float64 autograd oracle and against the fp32 SIMT execution of the same recurrences."""
    from lip_b200 import ggn
    hidden, n_out, in_dim, M, N = TC_ZCONFIGS[name]
    ost, lst = make_pair("large", hidden=hidden, n_out=n_out, in_dim=in_dim, seed=77)
    rng = np.random.default_rng(79)
    Z = rng.random((M, in_dim)).astype(np.float32)
    D = ost.flat()[0].size
    U, V = _probes(D, 2, 80)
    Y = rng.standard_normal((2, M, n_out)).astype(np.float32)
    ref = O.ggn_vp_zgrad(ost, Z, "classifier", U, V, full_set_size=N, per_probe=True)
    tc = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N, tensor_path=True)
    assert "tcgen05" in tc._lip_model.path_name()
    simt = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N, tensor_path=False)
    g_tc = tc.zgrad(cu(U), cu(V), per_probe=True).cpu().numpy()
    g_simt = simt.zgrad(cu(U), cu(V), per_probe=True).cpu().numpy()
    e_tc, e_simt = rel_err(g_tc, ref), rel_err(g_simt, ref)
    print(f"{name}: zgrad GGN tc rel err {e_tc:.2e}, simt rel err {e_simt:.2e}")
    assert e_simt < TOL and e_tc < TOL
    assert rel_err(tc.zgrad(cu(U), cu(V)).cpu().numpy(), ref.sum(0)) < TOL
    Wz_ref, WTz_ref = O.W_vps_zgrad(ost, Z, "classifier", full_set_size=N)
    Wf, WTf = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=N, tensor_path=True)
    assert rel_err(Wf.zgrad(cu(U), cu(Y)).cpu().numpy(), Wz_ref(U, Y)) < TOL
    assert rel_err(WTf.zgrad(cu(Y), cu(V)).cpu().numpy(), WTz_ref(Y, V)) < TOL


def test_zgrad_gelu_model_on_the_tensor_path_falls_back_to_simt():
    """A GELU regressor wide enough for the tcgen05 path: phi''/phi' is unbounded at GELU's stationary point, so lip_zgrad keeps dh and
    runs the fp32 SIMT recurrences even though the model's operators are on the tensor cores.  Single probe (B = 1) as well."""
    from lip_b200 import ggn
    ost, lst = make_pair("regressor", hidden=[128, 128], n_out=1, in_dim=96, seed=88, logvar=0.2)
    rng = np.random.default_rng(89)
    Z = rng.standard_normal((70, 96)).astype(np.float32)
    D = ost.flat()[0].size
    U, V = _probes(D, 1, 90)
    vp = ggn.compute_ggn_vp(lst, cu(Z), "regressor", full_set_size=700)
    assert "tcgen05" in vp._lip_model.path_name()
    ref = O.ggn_vp_zgrad(ost, Z, "regressor", U, V, full_set_size=700)
    assert rel_err(vp.zgrad(cu(U[0]), cu(V[0])).cpu().numpy(), ref) < TOL


def test_zgrad_is_cuda_graph_capturable():
    """lip_zgrad keeps the hot-call contract of include/lip_b200.h (no allocation / host sync inside the library; phi'' and phi''/phi'
    are bind-time state): one call captured into a CUDA graph replays to the same result on both executions."""
    from lip_b200 import ggn
    hidden, n_out, in_dim, M, N = TC_ZCONFIGS["tc_ragged"]
    ost, lst = make_pair("large", hidden=hidden, n_out=n_out, in_dim=in_dim, seed=77)
    rng = np.random.default_rng(91)
    Z = cu(rng.random((M, in_dim)).astype(np.float32))
    D = ost.flat()[0].size
    for tp in (False, True):
        vp = ggn.compute_ggn_vp(lst, Z, "classifier", full_set_size=N, tensor_path=tp)
        U, V = (cu(x) for x in _probes(D, 3, 92))
        eager = vp.zgrad(U, V).clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            vp.zgrad(U, V)                              # warm-up on the capture stream (workspace growth)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = vp.zgrad(U, V)
        graph.replay()
        torch.cuda.synchronize()
        assert rel_err(out.cpu().numpy(), eager.cpu().numpy()) < 1e-6
        V2 = cu(_probes(D, 3, 93)[1])
        expect = vp.zgrad(U, V2).clone()
        V.copy_(V2)
        graph.replay()
        torch.cuda.synchronize()
        assert rel_err(out.cpu().numpy(), expect.cpu().numpy()) < 1e-6


def test_jvp_zgrad_matches_oracle():
    from lip_b200 import _cabi, ggn
    ost, lst, Z, mt, N = _setup("mlp_ragged")
    D = ost.flat()[0].size
    _, V = _probes(D, 2, 14)
    Cb = np.random.default_rng(15).standard_normal((2, Z.shape[0], 7)).astype(np.float32)
    bm = ggn.compute_ggn_vp(lst, cu(Z), mt)._lip_model
    got = bm.zgrad(_cabi.ZGRAD_JVP, cu(V), cu(Cb), scale=1.0).cpu().numpy()
    assert rel_err(got, O.jvp_zgrad(ost, Z, Cb, V)) < TOL


def test_trace_estimator_gradient():
    """d/dZ of the Hutchinson estimate mean_b eps_b^T GGN(Z) eps_b (stochtrace.py:22-34 differentiated as the reference's
    objective differentiates its estimators): one lip_zgrad call with ubar = v = the probe block."""
    from lip_b200 import ggn
    ost, lst, Z, mt, N = _setup("C2_xor")
    D = ost.flat()[0].size
    E = np.random.default_rng(16).choice([-1.0, 1.0], size=(16, D)).astype(np.float32)
    ref = O.ggn_vp_zgrad(ost, Z, mt, E, E, full_set_size=N) / 16
    vp = ggn.compute_ggn_vp(lst, cu(Z), mt, full_set_size=N)
    got = vp.zgrad(cu(E), cu(E)).cpu().numpy() / 16
    assert rel_err(got, ref) < TOL


def test_zgrad_headline_shape_runs_and_matches_finite_difference():
    """C3b shape (784-1024-512-256-128-10, M = 512) is too large for the autograd oracle; check the directional derivative
    of s(Z) = u^T GGN(Z) v along a random direction against a central difference of the CUDA product itself."""
    from lip_b200 import ggn
    ost, lst = make_pair("large", hidden=[1024, 512, 256, 128], n_out=10, in_dim=784, seed=3)
    rng = np.random.default_rng(17)
    Z = rng.random((512, 784)).astype(np.float32)
    D = ost.flat()[0].size
    U, V = _probes(D, 2, 18)
    U *= 1e-2
    V *= 1e-2
    vp = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=60000, tensor_path=False)
    g = vp.zgrad(cu(U), cu(V)).double().cpu().numpy()
    g_tc = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=60000, tensor_path=True).zgrad(cu(U), cu(V)).double().cpu().numpy()
    assert rel_err(g_tc, g) < 2e-5          # tcgen05 3xTF32 execution against the fp32 SIMT execution at the headline shape
    dZ = rng.standard_normal(Z.shape)
    eps = 1e-2

    def s(Zv):
        f = ggn.compute_ggn_vp(lst, cu(Zv.astype(np.float32)), "classifier", full_set_size=60000, tensor_path=False)
        return float((cu(U).double() * f(cu(V)).double()).sum())

    fd = (s(Z + eps * dZ) - s(Z - eps * dZ)) / (2 * eps)
    an = float((g * dZ).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1e-12), (fd, an)


def test_lenet5_zgrad_matches_oracle():
    """lip_zgrad for relu conv stage programs (LeNet5, src/scalemodels.py:11-49): d/dZ with respect to the input IMAGES."""
    from lip_b200 import ggn
    ost, lst = make_pair("lenet5", seed=1)
    rng = np.random.default_rng(2)
    Z = rng.random((3, 28, 28, 1)).astype(np.float32)
    D = ost.flat()[0].size
    U, V = _probes(D, 2, 19)
    Y = rng.standard_normal((2, 3, 10)).astype(np.float32)
    ref = O.ggn_vp_zgrad(ost, Z, "classifier", U, V, full_set_size=600, per_probe=True)
    vp = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=600)
    got = vp.zgrad(cu(U), cu(V), per_probe=True).cpu().numpy().reshape(ref.shape)
    assert rel_err(got, ref) < TOL
    assert rel_err(vp.zgrad(cu(U), cu(V)).cpu().numpy().reshape(Z.shape), ref.sum(0)) < TOL
    Wz_ref, WTz_ref = O.W_vps_zgrad(ost, Z, "classifier", full_set_size=600)
    Wf, WTf = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=600)
    assert rel_err(Wf.zgrad(cu(U), cu(Y)).cpu().numpy().reshape(Z.shape), Wz_ref(U, Y)) < TOL
    assert rel_err(WTf.zgrad(cu(Y), cu(V)).cpu().numpy().reshape(Z.shape), WTz_ref(Y, V)) < TOL


@pytest.mark.parametrize("shape,M", [((8, 8, 3), 3), ((32, 32, 3), 2)])
def test_resnet1m_zgrad_matches_oracle(shape, M):
    """lip_zgrad for residual programs (ResNet1M, src/scalemodels.py:70-157; round 2): d/dZ with respect to the input IMAGES through
    convs, BatchNorm in eval mode (its scale tangent multiplies xhat(Z): the extra term of the q recurrence), skip connections, the 1x1
    strided shortcuts, the global mean and the head — against torch.func.grad through the oracle's float64 restatement of ggn.py."""
    from lip_b200 import ggn
    ost, lst = make_pair("resnet1m", n_out=10, in_shape=shape, seed=3)
    rng = np.random.default_rng(4)
    Z = rng.random((M,) + shape).astype(np.float32)
    D = ost.flat()[0].size
    U, V = _probes(D, 2, 23)
    Y = rng.standard_normal((2, M, 10)).astype(np.float32)
    ref = O.ggn_vp_zgrad(ost, Z, "classifier", U, V, full_set_size=500, per_probe=True)
    vp = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=500)
    got = vp.zgrad(cu(U), cu(V), per_probe=True).cpu().numpy().reshape(ref.shape)
    e = rel_err(got, ref)
    print(f"ResNet1M {shape} M={M}: ggn zgrad rel err {e:.2e} ({vp._lip_model.path_name()})")
    assert e < TOL
    assert rel_err(vp.zgrad(cu(U), cu(V)).cpu().numpy().reshape(Z.shape), ref.sum(0)) < TOL
    Wz_ref, WTz_ref = O.W_vps_zgrad(ost, Z, "classifier", full_set_size=500)
    Wf, WTf = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=500)
    assert rel_err(Wf.zgrad(cu(U), cu(Y)).cpu().numpy().reshape(Z.shape), Wz_ref(U, Y)) < TOL
    assert rel_err(WTf.zgrad(cu(Y), cu(V)).cpu().numpy().reshape(Z.shape), WTz_ref(Y, V)) < TOL


def test_zgrad_rejects_bad_shapes():
    from lip_b200 import ggn
    ost, lst, Z2, mt, N = _setup("C2_xor")
    vp2 = ggn.compute_ggn_vp(lst, cu(Z2), mt)
    with pytest.raises(ValueError):
        vp2.zgrad(torch.zeros(2, 5, device="cuda"), torch.zeros(2, 5, device="cuda"))


@pytest.mark.parametrize("name", ["C1_toy_sine", "C2_xor", "mlp_ragged"])
def test_build_WTWz_native_matches_oracle(name):
    """ggn.py:233-272: cross-Gram W_X^T W_Z through lip_gram_cross (two point sets, one layer program)."""
    from lip_b200 import ggn
    ost, lst, Z, mt, N = _setup(name)
    X = np.random.default_rng(21).standard_normal((Z.shape[0] + 5, Z.shape[1])).astype(np.float32)
    Wz_o, WzT_o = O.compute_W_vps(ost, Z, mt)
    W_o, WT_o = O.compute_W_vps(ost, X, mt)
    K = CONFIGS[name][2]
    dz, dx = Z.shape[0] * K, X.shape[0] * K
    shape_z = (Z.shape[0],) if mt == "regressor" else (Z.shape[0], K)
    ref = np.stack([np.asarray(WT_o(Wz_o(e.reshape(shape_z)))).reshape(-1) for e in np.eye(dz)], axis=1)
    Wz, WzT = ggn.compute_W_vps(lst, cu(Z), mt)
    W, WT = ggn.compute_W_vps(lst, cu(X), mt)
    got = ggn.build_WTWz(WT, Wz, shape_z, d=dx, block=1)
    assert got.shape == (dx, dz)
    assert rel_err(got.cpu().numpy(), ref) < TOL
