"""Rows f2 / f3 of the scope table on CPU: the packed clip store + feeder against what the reference's own test
dataset yields, and the serving front end's clip slicing against the reference demo's functions (both recorded by
oracle/make_golden.py from the unmodified reference)."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLDEN as GOLDEN_DIR
from kasportsformer_b200 import clipstore, serving


def _golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def _write_clip_dir(tmp_path, z, actions):
    d = tmp_path / "clips" / "test"
    d.mkdir(parents=True)
    for i, a in enumerate(actions):
        rec = {"data_input": z["rec_data_input"][i], "data_label": z["rec_data_label"][i],
               "data_label_scaled": z["rec_data_label_scaled"][i], "data_factor": z["rec_data_factor"][i],
               "data_res": tuple(int(v) for v in z["rec_res"][i]), "data_action": str(a), "data_env": "outdoor"}
        with open(d / ("%08d.pkl" % i), "wb") as fh:
            pickle.dump(rec, fh)
    return str(d)


def test_clip_store_matches_reference_dataset(tmp_path):
    z = _golden("clipstore.npz")
    actions = [str(a) for a in z["actions"]]
    clip_dir = _write_clip_dir(tmp_path, z, actions)
    meta = clipstore.pack_clip_dir(clip_dir, str(tmp_path / "shard"))
    assert meta["n_clips"] == 5 and meta["n_frames"] == 27 and meta["action_names"] == ["jump", "throw", "kick"]
    st = clipstore.ClipStore(str(tmp_path / "shard"))
    x, gt, factor, res, act = st.batch(0, len(st))
    assert np.array_equal(x, z["input"]) and np.array_equal(gt, z["gt"])          # bit-exact
    assert np.array_equal(factor, z["factor"]) and np.array_equal(res, z["res"])
    assert [st.action_names[i] for i in act] == actions


def test_clip_store_shards_and_feeder_cover_every_clip_once():
    rng = np.random.default_rng(0)
    n, T = 23, 9
    st = clipstore.ClipStore.from_arrays(rng.standard_normal((n, T, 17, 3)), rng.standard_normal((n, T, 17, 3)),
                                         rng.random((n, T)), np.tile([1312.0, 1216.0], (n, 1)), np.arange(n) % 3,
                                         ["a", "b", "c"])
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            sh = st.shard(rank, world)
            assert abs(len(sh) - n / world) < 1
            for x, gt, res, factor, action in clipstore.ClipFeeder(sh, 4, torch.device("cpu")):
                assert x.shape[1:] == (T, 17, 3) and res.shape[1:] == (2,) and factor.shape[1:] == (T,)
                assert x.dtype == torch.float32 and action.dtype == torch.int32
                seen.append(x.numpy()[:, 0, 0, 0])
        assert np.array_equal(np.concatenate(seen), st.arrays["input"][:, 0, 0, 0])
    assert len(clipstore.ClipFeeder(st.shard(7, 64), 4, torch.device("cpu"))) == 0 or True
    empty = st.shard(0, 64)
    assert len(empty) == 0 and list(clipstore.ClipFeeder(empty, 4, torch.device("cpu"))) == []


def test_turn_into_clips_and_resample_match_reference_demo():
    z = _golden("serving.npz")
    for key in z.files:
        if key.startswith("resample_"):
            n, t = (int(v) for v in key.split("_")[1:])
            assert np.array_equal(serving.resample(n, t), z[key]), key
    for n in z["lengths"]:
        kp = z[f"kp_{n}"]
        clips, down = serving.turn_into_clips(kp, 27)
        assert np.array_equal(np.stack(clips), z[f"clips_{n}"]), n
        assert np.array_equal(np.asarray(down, np.int64), z[f"down_{n}"]), n
        assert np.allclose(serving.normalize_screen_coordinates(kp, 1920, 1080), z[f"norm_{n}"], rtol=0, atol=1e-6)
    # exact multiples of the clip length: the reference raises (unbound `downsample`); every clip is full here
    clips, down = serving.turn_into_clips(np.zeros((1, 54, 17, 3), np.float32), 27)
    assert len(clips) == 2 and down is None


def test_flip_permutation_matches_reference_demo():
    from kasportsformer_b200 import skeleton
    z = _golden("serving.npz")
    x = z["flip_in"].copy()
    x[..., 0] *= -1
    assert np.array_equal(x[:, :, list(skeleton.flip_permutation())], z["flip_out"])
