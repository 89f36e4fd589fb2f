"""Where the end-to-end step time goes (bench shape: 65,536 voices x 4,096 frames, biquad, 4 voice ranges).

Times, per step over the sustain phase of the render, the same loop with one more piece of the e2e path
enabled each time: rows only -> + device mix -> + mix D2H (async, two in flight) -> + note-off table H2D.
Run on a GPU box: `PYTHONPATH=. python tools/exp_e2e.py [steps] [modes, e.g. 0,1]`.
"""
import sys
import time

import numpy as np
import torch

import os
import pathlib

import synth2_b200._lib as _s2lib
if os.environ.get("S2_EXP_LIB"):                      # A/B builds of the library (experiments only)
    _s2lib.LIB_PATH = pathlib.Path(os.environ["S2_EXP_LIB"]).resolve()
import synth2_b200 as s2
from synth2_b200 import bankgen

V, T, SR = 65536, 4096, 48000
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 200
MODES = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 5, 6]
SKIP = 16          # blocks of attack/decay/mod-envelope transients before the timed sustain phase


def main():
    dev = torch.device("cuda", 0)
    voices = bankgen.make_bank(V, 2_880_000, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    stream = torch.cuda.current_stream()
    if os.environ.get("S2_OWN_STREAM"):          # a non-blocking stream of the caller's instead of the legacy default
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
    bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=0, stream=stream)
    bank.set_pipeline(4)
    ring = [torch.empty((V, T), device=dev, dtype=torch.float32) for _ in range(2)]
    bus_dev = torch.empty(T, device=dev, dtype=torch.float32)
    bus_host = [torch.empty(T, dtype=torch.float32).pin_memory() for _ in range(2)]
    rel_host = torch.from_numpy(voices["release_offset"].astype(np.uint32).view(np.int32)).pin_memory()
    state0 = bank.get_state()
    done = [torch.cuda.Event() for _ in range(2)]
    extra_ev = [torch.cuda.Event() for _ in range(4)]

    def run(mode):
        bank.set_state(state0)
        for i in range(SKIP):
            bank.render(T, ring[i & 1], T, None)
        bank.join(stream)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for i in range(STEPS):
            if mode >= 3:
                bank.set_releases(rel_host)
            if mode == 0:
                bank.render(T, ring[i & 1], T, None)
            elif mode == 1:
                bank.render(T, ring[i & 1], T, bus_dev)
            elif mode == 5:                      # device mix + an event on the caller's stream per step
                bank.render(T, ring[i & 1], T, bus_dev)
                done[i & 1].record(stream)
            elif mode == 9:                      # host-buffer call, nothing on the caller's stream (S2_EXP_NO_TAIL_WAIT=1)
                bank.set_releases(rel_host)
                bank.render_bus_host_async(T, bus_host[i & 1], ring[i & 1], T)
            elif mode in (7, 8):                 # rows only + 1 / 4 event records on the caller's stream per step
                bank.render(T, ring[i & 1], T, None)
                for k in range(1 if mode == 7 else 4):
                    extra_ev[k].record(stream)
            elif mode == 6:                      # device mix + join on the caller's stream per step
                bank.render(T, ring[i & 1], T, bus_dev)
                bank.join(stream)
            else:
                bank.render_bus_host_async(T, bus_host[i & 1], ring[i & 1], T)
                done[i & 1].record(stream)
                if i > 0 and mode != 4:
                    done[(i - 1) & 1].synchronize()
        enq = (time.perf_counter() - t0) * 1e3
        bank.join(stream)
        ev1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        return ev0.elapsed_time(ev1) / STEPS, wall / STEPS, enq / STEPS

    names = ["rows only", "+ device mix", "+ mix D2H, host waits for step i-1", "+ note-off table H2D",
             "same, host never waits (enqueue-bound?)", "device mix + event record on the caller's stream",
             "device mix + join on the caller's stream", "rows only + 1 event record per step", "rows only + 4 event records per step",
             "H2D + render + mix + D2H, nothing on the caller's stream"]
    # (the GPU slows down by a tenth over the first seconds of continuous rendering, profiles/r2_notes.md: warm it
    # up first and give the modes in an interleaved order, e.g. 0,1,2,3,0,1,2,3, or the later ones look worse)
    for _ in range(6):
        run(0)
    for mode in MODES:
        name = names[mode]
        dev_ms, wall_ms, enq_ms = run(mode)
        print(f"{name:45s} device {dev_ms*1e3:7.1f} us/step   wall {wall_ms*1e3:7.1f} us/step   host loop {enq_ms*1e3:7.1f} us/step", flush=True)
    bank.close()


if __name__ == "__main__":
    main()
