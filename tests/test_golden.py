"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit-for-bit.  GPU: the CUDA path matches them."""
import pathlib
import zlib

import numpy as np
import pytest

import oracle
from golden.make_golden import BANK_FIXTURES, EVENTS, SR, TOTAL, render_config1

BANK_NAMES = [name for _, name in BANK_FIXTURES]

G = pathlib.Path(__file__).resolve().parent / "golden"


def test_oracle_reproduces_config1_fixture():
    g = np.load(G / "synth_config1.npz")
    buf = render_config1()
    assert np.uint32(zlib.crc32(buf.tobytes())) == g["crc32"]
    for w, d in zip(g["windows"], g["window_data"]):
        assert buf[w:w + 64].tobytes() == d.tobytes()
    assert buf[::4801].tobytes() == g["stride_samples"].tobytes()
    assert float(np.abs(buf).max()) == float(g["peak"]) > 0.1


@pytest.mark.parametrize("name", BANK_NAMES)
def test_oracle_reproduces_bank_fixture(name):
    g = np.load(G / f"{name}.npz")
    v = g["voices"]
    st = oracle.bank_init_states(v)
    out, bus = oracle.bank_render(v, st, SR, int(g["filter_kind"]), 1000)
    assert out.tobytes() == g["out"].tobytes()
    assert bus.tobytes() == g["bus"].tobytes()
    assert st.tobytes() == g["state"].tobytes()
    assert np.all(out[5] == 0.0) and np.abs(out).max() > 0.05


@pytest.mark.gpu
def test_gpu_matches_config1_fixture():
    import synth2_b200 as s2
    g = np.load(G / "synth_config1.npz")
    syn = s2.Synth()
    buf = np.zeros(TOTAL, dtype=np.float32)
    cuts = sorted({0, TOTAL, *[f for f, _, _ in EVENTS]})
    for a, b in zip(cuts[:-1], cuts[1:]):
        for f, op, note in EVENTS:
            if f == a:
                syn.note_on(note, 1.0) if op == "on" else syn.note_off(note)
        syn.sample(buf[a:b], SR)
    syn.close()
    for w, d in zip(g["windows"], g["window_data"]):
        assert float(np.max(np.abs(buf[w:w + 64] - d))) <= 1e-4
    assert float(np.max(np.abs(buf[::4801] - g["stride_samples"]))) <= 1e-4
    assert abs(float((buf.astype(np.float64) ** 2).sum()) - float(g["sumsq"])) <= 1e-6 * float(g["sumsq"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", BANK_NAMES)
def test_gpu_matches_bank_fixture(name):
    import torch
    import synth2_b200 as s2
    g = np.load(G / f"{name}.npz")
    v = g["voices"]
    out = torch.zeros((16, 1000), device="cuda", dtype=torch.float32)
    bus = torch.zeros(1000, device="cuda", dtype=torch.float32)
    with s2.VoiceBank(v, SR, int(g["filter_kind"])) as bank:
        bank.render(1000, out, 1000, bus)
        bank.sync()
        st = bank.get_state()
    got = out.cpu().numpy()
    assert float(np.max(np.abs(got - g["out"]))) <= 1e-4
    assert st["phase"].tobytes() == g["state"]["phase"].tobytes()          # bit-exact
    assert np.array_equal(st["frame_offset"], g["state"]["frame_offset"])
    # 16 voices = one warp: the bus is summed in the reference's voice order
    mix = np.zeros(1000, np.float32)
    for r in got:
        mix = mix + r
    assert bus.cpu().numpy().tobytes() == mix.tobytes()
