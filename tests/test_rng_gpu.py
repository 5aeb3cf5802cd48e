"""The DEVICE generator of the image environments' autoreset payloads (csrc/apg_payload.cu) against numpy itself:
every value and the three Generator states after the payload must be bit-identical to the reference's draw order
(image_localization.py:153-156, dataset_iterator.py:57, image_perception_module.py:131-133, 159-161)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_payload(gens, Ng, M, loc, inv):
    g0, g1, g2 = gens
    out = {}
    if loc:
        out["target"] = g0.uniform(-1, 1, (Ng, 2)).astype(np.float32)
    out["idx"] = g2.integers(0, M, Ng)
    if inv:
        out["inv"] = (g1.integers(0, 2, size=Ng) == 1)
    out["pos"] = g1.uniform(-1, 1, (Ng, 2))
    return out


def _device_payload(states, Ng, off, N, M, loc, inv):
    from ap_gym_b200 import _capi

    lib = _capi.lib()
    idx = np.empty(N, np.int64)
    pos = np.empty((N, 2), np.float64)
    tgt = np.empty((N, 2), np.float32)
    iv = np.empty(N, np.uint8)
    _capi.check(lib.apg_rng_device_payload(0, C.byref(states[0]), C.byref(states[1]), C.byref(states[2]), Ng, off, N, M,
                                           int(loc), int(inv), idx.ctypes.data, pos.ctypes.data, tgt.ctypes.data,
                                           iv.ctypes.data))
    return {"idx": idx, "pos": pos, "target": tgt, "inv": iv.astype(bool)}


CASES = [
    # Ng, off, N, M, loc, inv, pre-draws (32-bit integers pulled from g1/g2 first, to leave a buffered half-word)
    (4096, 0, 4096, 60000, True, False, 0),
    (4096, 0, 4096, 10, False, True, 0),
    (4097, 0, 4097, 60000, True, True, 1),        # odd batch + buffered half-words: stream parity flips
    (5000, 1234, 777, 100000, True, True, 0),     # one shard of a larger global batch
    (4096, 0, 4096, 50_000_000, False, False, 0),  # ~1 % of the Lemire draws are rejected and shift the stream
    (3001, 1000, 2001, 33_554_431, True, True, 1),
    (64, 0, 64, 1, False, False, 0),              # len(dataset) == 1: integers() consumes nothing
    (1 << 17, 0, 1 << 17, 100000, True, False, 0),
]


@pytest.mark.parametrize("Ng,off,N,M,loc,inv,pre", CASES)
def test_device_payload_matches_numpy(Ng, off, N, M, loc, inv, pre):
    from ap_gym_b200 import _capi

    gens = [np.random.Generator(np.random.PCG64(s)) for s in (11, 22, 33)]
    for _ in range(pre):
        gens[1].integers(0, 7)
        gens[2].integers(0, 7)
    states = [_capi.Pcg64State.from_numpy(g) for g in gens]
    for rounds in range(3):  # consecutive payloads: the device-advanced states feed the next one
        want = _numpy_payload(gens, Ng, M, loc, inv)
        got = _device_payload(states, Ng, off, N, M, loc, inv)
        sl = slice(off, off + N)
        assert np.array_equal(got["idx"], want["idx"][sl]), f"payload {rounds}: bank indices"
        assert np.array_equal(got["pos"], want["pos"][sl]), f"payload {rounds}: positions"
        if loc:
            assert np.array_equal(got["target"], want["target"][sl]), f"payload {rounds}: targets"
        if inv:
            assert np.array_equal(got["inv"], want["inv"][sl]), f"payload {rounds}: label inversion"
        for k, (g, st) in enumerate(zip(gens, states)):
            ref = _capi.Pcg64State.from_numpy(g)
            for f, _ in _capi.Pcg64State._fields_:
                if f == "uinteger" and not ref.has_uint32:
                    continue  # stale half-word of an empty buffer
                assert getattr(st, f) == getattr(ref, f), f"payload {rounds}: stream g{k}.{f}"


def test_device_payload_refuses_rejection_heavy_banks():
    from ap_gym_b200 import _capi

    gens = [np.random.Generator(np.random.PCG64(s)) for s in (1, 2, 3)]
    states = [_capi.Pcg64State.from_numpy(g) for g in gens]
    with pytest.raises(Exception):
        _device_payload(states, 4096, 0, 4096, 3 * (1 << 30), False, False)  # 25 % of the draws are rejected


def test_engine_autoresets_identical_with_host_and_device_generators(monkeypatch):
    """Three episodes of a seeded vector env: the in-stream device payloads and the host-drawn ones give the same run."""
    import ap_gym_b200 as apg
    from oracle.datasets import synthetic_u8_bank

    imgs, labels = synthetic_u8_bank(seed=5, M=300, H=28, W=28, C=1, K=10)
    runs = []
    for host in ("1", "0"):
        monkeypatch.setenv("APG_IMG_HOST_RNG", host)
        cfg = apg.ImagePerceptionConfig(dataset=apg.ArrayImageDataset(imgs, labels, 10), step_limit=4)
        env = apg.ImageClassificationVectorEnv(257, cfg, output="numpy")
        env.reset(seed=3)
        rng = np.random.default_rng(0)
        rec = []
        for _ in range(16):
            a = rng.uniform(-1, 1, (257, 2)).astype(np.float32)
            p = rng.standard_normal((257, 10)).astype(np.float32)
            o, r, te, tr, info = env.step({"action": a, "prediction": p})
            rec.append((o["glimpse"].copy(), r.copy(), te.copy(), info["index"].copy()))
        st = [env.np_random.bit_generator.state] if hasattr(env, "np_random") else []
        runs.append((rec, st))
        env.close()
    for (g0, r0, t0, i0), (g1, r1, t1, i1) in zip(runs[0][0], runs[1][0]):
        assert np.array_equal(g0, g1) and np.array_equal(r0, r1) and np.array_equal(t0, t1) and np.array_equal(i0, i1)
    assert runs[0][1] == runs[1][1]
