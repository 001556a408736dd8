"""Shared builders: the same weights / points as an oracle state (float64 CPU) and as a lip_b200 state (CUDA)."""
import numpy as np

from oracle import models as OM


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def make_pair(kind, *, hidden=(), n_out=10, in_dim=0, seed=0, logvar=0.0, toy_layout=False, in_shape=None):
    """kind: 'regressor' | 'classifier' | 'large' | 'lenet5' | 'resnet1m'.  Returns (oracle_state, lip_state)."""
    import lip_b200  # noqa: F401
    from lip_b200 import scalemodels, toymodels

    if kind == "regressor":
        om = OM.OracleModel("regressor_mlp", (in_dim,), list(hidden), 1, "regressor")
        mod = toymodels.SimpleRegressor(hidden[0] if hidden else 1, len(hidden))
    elif kind == "classifier":
        om = OM.OracleModel("classifier_mlp", (in_dim,), list(hidden), n_out, "classifier")
        mod = toymodels.SimpleClassifier(hidden[0] if hidden else 1, len(hidden), n_out)
    elif kind == "large":
        shp = tuple(in_shape) if in_shape else (in_dim,)
        om = OM.OracleModel("large_classifier", shp, list(hidden), n_out, "classifier")
        mod = scalemodels.LargeClassifier(shp, list(hidden), len(hidden), n_out)
    elif kind == "lenet5":
        om = OM.LeNet5()
        mod = scalemodels.LeNet5()
    elif kind == "resnet1m":
        om = OM.ResNet1M(n_out, tuple(in_shape) if in_shape else (32, 32, 3))
        mod = scalemodels.ResNet1M(n_out)
    else:
        raise ValueError(kind)
    variables = om.init(seed)
    ost = OM.OracleState(om, variables, toy_layout=toy_layout, logvar=logvar)
    lst = scalemodels.TrainState(params=ost.params, apply_fn=mod.apply, batch_stats=ost.batch_stats)
    return ost, lst
