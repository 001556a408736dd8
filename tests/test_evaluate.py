"""the evaluation consumer of predict_lla_scalable (scale_experiments/evaluate.py:40-231).
CPU: the oracle's metrics against hand-computed cases.  GPU: lip_mc_softmax_predictive and batch_nll / eval_dataset_extended
through the CUDA path against the oracle with identical posterior noise."""
import math

import numpy as np
import pytest

from helpers import make_pair, rel_err
from oracle import lip_oracle as O


def test_oracle_metrics_known_answers():
    probs = np.array([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1], [0.3, 0.3, 0.4], [0.05, 0.05, 0.9]])
    labels = np.array([0, 1, 0, 2])
    # Brier: mean of squared distances to the one-hot label
    expect = np.mean([0.09 + 0.04 + 0.01, 0.01 + 0.04 + 0.01, 0.49 + 0.09 + 0.16, 0.0025 + 0.0025 + 0.01])
    assert abs(O.brier_score(probs, labels) - expect) < 1e-12
    # ECE with 2 bins: confidences 0.7, 0.8, 0.4, 0.9 -> bin [0,.5): {0.4, wrong}; bin [.5,1): {0.7, 0.8, 0.9 all right}
    expect = abs(0.4 - 0.0) * 0.25 + abs(0.8 - 1.0) * 0.75
    assert abs(O.ece(probs, labels, n_bins=2) - expect) < 1e-12
    # one sample: the MC average is the softmax itself
    logits = np.log(probs)[None]
    la, mean = O.mc_softmax_predictive(logits, labels)
    np.testing.assert_allclose(mean, probs, rtol=1e-12)
    np.testing.assert_allclose(la, np.log(probs[np.arange(4), labels]), rtol=1e-12)
    # two samples: log of the averaged true-class probability
    l2 = np.stack([np.log(probs), np.log(probs[::-1])])
    la, mean = O.mc_softmax_predictive(l2, labels)
    pt = 0.5 * (probs[np.arange(4), labels] + probs[::-1][np.arange(4), labels])
    np.testing.assert_allclose(la, np.log(pt), rtol=1e-12)


def test_package_metrics_match_oracle_on_cpu_arrays():
    """brier_score / ece / ood_scores are host-side numpy in the reference and here: no GPU needed."""
    import importlib
    rng = np.random.default_rng(3)
    p = rng.dirichlet(np.ones(5), size=200)
    y = rng.integers(0, 5, size=200)
    ev = importlib.import_module("lip_b200.evaluate")
    assert abs(ev.brier_score(p, y) - O.brier_score(p, y)) < 1e-12
    assert abs(ev.ece(p, y) - O.ece(p, y)) < 1e-12
    np.testing.assert_array_equal(ev.ood_scores(p), -p.max(1))


@pytest.mark.gpu
@pytest.mark.parametrize("S,B,C", [(1, 1, 2), (7, 33, 10), (300, 257, 10), (5, 1000, 64)])
def test_mc_softmax_kernel_matches_oracle(S, B, C):
    import torch
    from lip_b200 import evaluate as ev
    rng = np.random.default_rng(S * 1000 + B)
    logits = (4.0 * rng.standard_normal((S, B, C))).astype(np.float32)
    logits[0, 0, :] += 80.0                      # large offsets must not overflow
    y = rng.integers(0, C, size=B)
    la_ref, mean_ref = O.mc_softmax_predictive(logits, y)
    la, mean = ev.mc_softmax_predictive(torch.as_tensor(logits, device="cuda"), torch.as_tensor(y, device="cuda"))
    assert rel_err(mean.cpu().numpy(), mean_ref) < 1e-5
    np.testing.assert_allclose(la.cpu().numpy(), la_ref, rtol=1e-4, atol=1e-5)
    _, mean2 = ev.mc_softmax_predictive(torch.as_tensor(logits, device="cuda"))           # no labels: OOD pass
    assert rel_err(mean2.cpu().numpy(), mean_ref) < 1e-5
    with pytest.raises(ValueError):
        ev.mc_softmax_predictive(torch.zeros(2, 3, 65, device="cuda"))


@pytest.mark.gpu
def test_batch_nll_and_eval_dataset_match_oracle():
    import torch
    from lip_b200 import evaluate as ev
    from test_gpu_parity import _setup
    ost, lst, Z, _, _ = _setup("C2_xor")          # the configuration test_sampler_and_predictive_match_oracle pins the sampler on
    Z = Z[:12]
    rng = np.random.default_rng(42)
    D = ost.flat()[0].size
    alpha, N, S = 2.5, 800, 6
    batches = [(rng.standard_normal((9, 2)).astype(np.float32), rng.integers(0, 2, size=(9, 1))),
               (rng.standard_normal((5, 2)).astype(np.float32), rng.integers(0, 2, size=(5, 1)))]
    Eps = rng.standard_normal((S, D)).astype(np.float32)
    cu = lambda a: torch.as_tensor(a, device="cuda")
    x, y = batches[0]
    nll_ref, acc_ref, mean_ref = O.batch_nll(ost, x, y, Z, alpha=alpha, full_set_size=N, model_type="classifier", Eps=Eps.astype(np.float64))
    nll, acc, mean = ev.batch_nll(lst, cu(x), cu(y), cu(Z), alpha=alpha, full_set_size=N, model_type="classifier", num_mc_samples=S,
                                  rng=0, return_mean=True, eps=cu(Eps))
    assert abs(float(nll) - nll_ref) <= 1e-4 * abs(nll_ref)
    assert abs(float(acc) - acc_ref) < 1e-6
    err = rel_err(mean.cpu().numpy(), mean_ref)
    print(f"batch_nll: nll {float(nll):.6f} vs {nll_ref:.6f}, mean-prob rel err {err:.2e}")
    assert err < 1e-4
    # the cached sampler (built once per dataset) gives the same numbers as the per-batch rebuild
    sampler = ev._sampler_for(lst, cu(Z), alpha, "classifier", N)
    nll2, _ = ev.batch_nll(lst, cu(x), cu(y), cu(Z), alpha=alpha, full_set_size=N, model_type="classifier", num_mc_samples=S, rng=0,
                           eps=cu(Eps), sampler=sampler)
    assert abs(float(nll2) - float(nll)) <= 1e-6 * abs(float(nll))
    with pytest.raises(ValueError):
        ev.batch_nll(lst, cu(x), cu(y), cu(Z), alpha=alpha, full_set_size=N, model_type="classifier", num_mc_samples=S, rng=0, scalable=False)
    # dataset loop: shapes, ranges and consistency of the aggregate with its own per-batch pieces
    tot_nll, tot_acc, bri, cal, probs, labels = ev.eval_dataset_extended(lst, [(cu(a), cu(b)) for a, b in batches], cu(Z), alpha, N,
                                                                        "classifier", 64, rng=7)
    assert probs.shape == (14, 2) and labels.shape == (14,)
    np.testing.assert_allclose(probs.sum(1), 1.0, rtol=1e-5)
    assert abs(bri - O.brier_score(probs, labels)) < 1e-9 and abs(cal - O.ece(probs, labels)) < 1e-9
    assert 0.0 <= tot_acc <= 1.0 and math.isfinite(tot_nll)
    nll_d, acc_d = ev.eval_dataset(lst, [(cu(a), cu(b)) for a, b in batches], cu(Z), alpha, N, "classifier", 64, rng=7)
    assert abs(nll_d - tot_nll) < 1e-6 and abs(acc_d - tot_acc) < 1e-9          # same keys -> same samples
    auc = ev.auroc_ood(lst, probs, [(cu(10.0 + batches[1][0]), None)], cu(Z), alpha, N, "classifier", 16, rng=3)
    assert 0.0 <= auc <= 1.0


@pytest.mark.gpu
def test_numpy_batch_generator_never_hits_a_stale_bound_model():
    """the bind cache used to key non-torch inputs on id(Z); CPython recycles ids, so a generator that yields
    numpy batches alternates between two ids and batch i+2 was scored with the logits of batch i.  Only torch tensors are cached
    now (identity + version counter); numpy batches are bound afresh.  Checked on the model outputs and through eval_dataset."""
    import torch
    from lip_b200 import evaluate as ev, ggn
    from test_gpu_parity import _setup
    ost, lst, Z, _, _ = _setup("C2_xor")
    Z = Z[:12]
    rng = np.random.default_rng(5)
    data = [rng.standard_normal((7, 2)).astype(np.float32) * (1 + i) for i in range(6)]

    def gen():
        for a in data:
            yield a.copy()                   # a fresh temporary every time: ids recycle

    for i, xb in enumerate(gen()):
        bm = ggn._bind(lst, xb, "classifier")
        assert rel_err(bm.outputs().cpu().numpy(), O.model_outputs(ost, data[i])) < 2e-6, i
    # in-place edits of a torch tensor invalidate its entry
    zt = torch.as_tensor(data[0], device="cuda").clone()
    o1 = ggn._bind(lst, zt, "classifier").outputs().clone()
    assert ggn._bind(lst, zt, "classifier") is ggn._bind(lst, zt, "classifier")
    zt.mul_(3.0)
    o2 = ggn._bind(lst, zt, "classifier").outputs()
    assert rel_err(o2.cpu().numpy(), O.model_outputs(ost, 3.0 * data[0])) < 2e-6 and not torch.allclose(o1, o2)
    # the dataset loop over numpy batches: per-batch NLL pieces equal the ones computed from CUDA tensors held in a list
    labels = [rng.integers(0, 2, size=(7, 1)) for _ in data]
    cu = lambda a: torch.as_tensor(a, device="cuda")
    nll_np, acc_np = ev.eval_dataset(lst, ((a.copy(), b) for a, b in zip(data, labels)), cu(Z), 2.5, 800, "classifier", 32, rng=7)
    nll_t, acc_t = ev.eval_dataset(lst, [(cu(a), cu(b)) for a, b in zip(data, labels)], cu(Z), 2.5, 800, "classifier", 32, rng=7)
    assert abs(nll_np - nll_t) <= 1e-6 * abs(nll_t) and abs(acc_np - acc_t) < 1e-9
