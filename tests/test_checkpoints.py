"""checkpoint formats (src/utils.py:20-75) — .npy inducing points and the flax msgpack state file.  flax is absent;
the layout is restated from its published serialization format (PARITY UNPINNED, see utils.py) and pinned here by a hand-assembled
byte string and round trips.  CPU only."""
import struct

import numpy as np
import pytest


def _utils():
    import importlib
    return importlib.import_module("lip_b200.utils")


def test_hand_assembled_flax_msgpack_bytes_decode():
    U = _utils()
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    raw = a.tobytes()
    # ext payload = msgpack array of 3: [shape array, dtype str, bin]
    payload = b"\x93" + b"\x92\x02\x03" + b"\xa7float32" + b"\xc4" + bytes([len(raw)]) + raw
    ext = b"\xc7" + bytes([len(payload)]) + b"\x01" + payload                     # ext8, type 1 (ndarray)
    blob = b"\x81" + b"\xa6params" + b"\x81" + b"\xa6kernel" + ext                # {"params": {"kernel": <ndarray>}}
    sd = U.state_dict_from_bytes(blob)
    np.testing.assert_array_equal(sd["params"]["kernel"], a)
    assert sd["params"]["kernel"].dtype == np.float32
    # and our writer produces bytes the same decoder reads back, with ExtType 1 leaves
    out = U.state_dict_to_bytes({"params": {"kernel": a}})
    assert out[:1] == b"\x81" and b"\xa7float32" in out
    np.testing.assert_array_equal(U.state_dict_from_bytes(out)["params"]["kernel"], a)
    # numpy scalars (ExtType 3)
    sc = U.state_dict_from_bytes(U.state_dict_to_bytes({"step": np.int32(7)}))
    assert sc["step"] == 7


def test_model_checkpoint_round_trip_latest_step(tmp_path):
    import torch
    U = _utils()
    from lip_b200 import scalemodels
    rng = np.random.default_rng(0)
    params = {"Dense_0": {"bias": rng.standard_normal(4).astype(np.float32), "kernel": rng.standard_normal((3, 4)).astype(np.float32)},
              "Dense_1": {"bias": torch.zeros(2), "kernel": torch.ones(4, 2)}}
    st = scalemodels.TrainState(params=params, apply_fn=None, batch_stats={"BatchNorm_0": {"mean": np.zeros(4, np.float32), "var": np.ones(4, np.float32)}})
    U.save_checkpoint(st, tmp_path, "map_mlp", 3)
    p2 = {k: {kk: np.asarray(vv) * 2 for kk, vv in v.items()} for k, v in params.items()}
    U.save_checkpoint(scalemodels.TrainState(params=p2, apply_fn=None, batch_stats=st.batch_stats), tmp_path, "map_mlp", 12)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["map_mlp_12"]           # overwrite=True / keep=1
    got = U.load_checkpoint(tmp_path, "map_mlp", target=st)
    np.testing.assert_array_equal(got.params["Dense_0"]["kernel"], 2 * params["Dense_0"]["kernel"])
    np.testing.assert_array_equal(got.batch_stats["BatchNorm_0"]["var"], np.ones(4, np.float32))
    raw = U.load_checkpoint(tmp_path, "map_mlp")                                  # target=None: the raw state dict
    assert raw["step"] == 12 and set(raw["params"]) == {"Dense_0", "Dense_1"}
    assert U.load_checkpoint(tmp_path, "other", target=st) is st                  # nothing found: the target comes back
    bad = scalemodels.TrainState(params={"Dense_0": params["Dense_0"]}, apply_fn=None)
    with pytest.raises(ValueError):
        U.load_checkpoint(tmp_path, "map_mlp", target=bad)
    assert U.count_model_params(got.params) == 4 + 12 + 2 + 8


def test_array_checkpoint_round_trip(tmp_path):
    U = _utils()
    Z = np.random.default_rng(1).random((5, 28, 28, 1)).astype(np.float32)
    fn = U.save_array_checkpoint(Z, tmp_path, "ip_mnist", 40)
    assert fn.endswith("ip_mnist_40.npy")
    np.testing.assert_array_equal(U.load_array_checkpoint(tmp_path, "ip_mnist", 40, device="cpu"), Z)
    with pytest.raises(FileNotFoundError):
        U.load_array_checkpoint(tmp_path, "ip_mnist", 41, device="cpu")


def test_full_train_state_with_opt_state_round_trips(tmp_path):
    """the reference saves the WHOLE flax TrainState (utils.py:46-60), opt_state included; flax's from_state_dict with
    target=TrainState raises on a file without it.  A state shaped like optax's adam (a tuple of (ScaleByAdamState(count, mu, nu),
    EmptyState())) is written in flax's index-keyed layout and restored."""
    import collections
    import dataclasses
    U = _utils()
    rng = np.random.default_rng(2)
    params = {"Dense_0": {"bias": rng.standard_normal(3).astype(np.float32), "kernel": rng.standard_normal((2, 3)).astype(np.float32)}}
    Adam = collections.namedtuple("ScaleByAdamState", ["count", "mu", "nu"])
    zeros = {"Dense_0": {"bias": np.zeros(3, np.float32), "kernel": np.zeros((2, 3), np.float32)}}
    opt = (Adam(np.int32(5), {k: {kk: vv + 1 for kk, vv in v.items()} for k, v in zeros.items()}, zeros), ())

    @dataclasses.dataclass
    class FullState:
        params: dict
        opt_state: tuple
        batch_stats: dict
        step: int = 0

    st = FullState(params=params, opt_state=opt, batch_stats={})
    U.save_checkpoint(st, tmp_path, "map_full", 9)
    raw = U.load_checkpoint(tmp_path, "map_full")
    assert set(raw) == {"step", "params", "opt_state", "batch_stats"}
    assert set(raw["opt_state"]) == {"0", "1"} and set(raw["opt_state"]["0"]) == {"count", "mu", "nu"}
    assert raw["opt_state"]["0"]["count"] == 5
    np.testing.assert_array_equal(raw["opt_state"]["0"]["mu"]["Dense_0"]["bias"], np.ones(3, np.float32))
    got = U.load_checkpoint(tmp_path, "map_full", target=st)
    np.testing.assert_array_equal(got.params["Dense_0"]["kernel"], params["Dense_0"]["kernel"])
    np.testing.assert_array_equal(got.opt_state["0"]["mu"]["Dense_0"]["kernel"], np.ones((2, 3), np.float32))
