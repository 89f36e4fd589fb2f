#!/usr/bin/env python3
"""bench.py — voice-samples/sec of the render hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 3, per GPU): 65,536 voices, oscillator + resonant biquad (2nd-order
low-pass) + ADSR, 48 kHz, 60 s = 2,880,000 frames per voice, streamed as 4,096-frame blocks through
a ring of two 1 GiB voice-major output buffers with the DSP state carried on the device.
One "step" = one block of every voice (the last of the default 704 steps is the 512-frame
remainder).  N GPUs = N x 65,536 voices, each rank owning a contiguous voice range ("weak").

  value      device-resident throughput: inputs (voice table, state) already in HBM, CUDA events
             around the K steps on the launching stream, max over ranks.
  e2e        the same render driven through the host-buffer C-ABI calls a streaming caller uses:
             every step uploads the note-off table (pinned host -> device), renders the block
             (per-voice output stays in the device ring) with the mono mix, and copies the mix
             (the `Synth::sample` result) back to pinned host memory.
  roofline   HBM: 4 algorithmic bytes per voice-sample (one f32 store, touched once)
  cpu_baseline / --impl reference
             the CPU port of the reference path (oracle/, "port": the Rust reference cannot be
             built here) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import pathlib
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR = 48000
VOICES_PER_GPU = 65536
BLOCK = 4096
RENDER_FRAMES = 60 * SR          # 2,880,000
DEFAULT_STEPS = (RENDER_FRAMES + BLOCK - 1) // BLOCK   # 704
FILTER_BIQUAD = 1
BYTES_PER_VOICE_SAMPLE = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=DEFAULT_STEPS)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voices", type=int, default=VOICES_PER_GPU, help="voices per GPU")
    ap.add_argument("--block", type=int, default=BLOCK)
    ap.add_argument("--master-bus", action="store_true",
                    help="N>1: also mix every rank's voices and NCCL-reduce the whole-render bus to rank 0 "
                         "(BASELINE config 4), inside the timed region")
    ap.add_argument("--pipeline", type=int, default=4,
                    help="render each GPU's bank as this many voice ranges on internal streams (1 = single stream)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=30.0, help="target CPU time of the baseline sample")
    return ap.parse_args()


def step_frames(steps, block):
    """Frames of each step: 4,096 each; the default 704 steps end exactly at 60 s."""
    if block == BLOCK and steps == DEFAULT_STEPS:
        return [BLOCK] * (steps - 1) + [RENDER_FRAMES - BLOCK * (steps - 1)]
    return [block] * steps


# ------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            "nvmlClocksEventReasonHwSlowdown": "hw_slowdown",
            "nvmlClocksThrottleReasonHwSlowdown": "hw_slowdown",
            "nvmlClocksThrottleReasonHwThermalSlowdown": "hw_thermal_slowdown",
            "nvmlClocksThrottleReasonSwThermalSlowdown": "sw_thermal_slowdown",
            "nvmlClocksThrottleReasonSwPowerCap": "sw_power_cap",
            "nvmlClocksThrottleReasonHwPowerBrakeSlowdown": "hw_power_brake",
        }
        masks = {getattr(nv, k): v for k, v in names.items() if hasattr(nv, k)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for m, name in masks.items():
                    if r & m:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self._nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------
# CPU arm: the port of the reference path (oracle/) on the host cores

_ORACLE_FLAGS = None


def cpu_oracle():
    """The C port of the reference path, built for this host (oracle/Makefile `native`)."""
    global _ORACLE_FLAGS
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle
    if _ORACLE_FLAGS is None:
        _ORACLE_FLAGS = oracle.use_native_build()
    return oracle


def cpu_port_throughput(voices_desc, n_threads, frames, repeats=1):
    oracle = cpu_oracle()
    best = None
    for _ in range(repeats):
        st = oracle.bank_init_states(voices_desc)
        t0 = time.perf_counter()
        # per-thread scratch rows instead of a [V][frames] matrix: the sample can then be long enough
        oracle.bank_render(voices_desc, st, SR, FILTER_BIQUAD, frames, want_voices=False, want_bus=False,
                           nthreads=n_threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return voices_desc.shape[0] * frames / best, best


def cpu_baseline(args, bankgen):
    cores = os.cpu_count() or 1
    probe = bankgen.make_bank(max(cores * 4, 32), RENDER_FRAMES, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    rate, _ = cpu_port_throughput(probe, cores, 2048)
    want = max(rate * args.cpu_seconds, 1.0)
    nv = max(cores * 64, 1024)
    frames = int(min(max(want / nv // BLOCK, 1), 720)) * BLOCK      # at most one 60 s render per voice
    sample = bankgen.make_bank(nv, RENDER_FRAMES, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    rate, dt = cpu_port_throughput(sample, cores, frames)
    return {"value": rate, "unit": "voice-samples/s", "cores": cores, "kind": "port",
            "sample": f"{nv} voices x {frames} frames of the config-3 bank (biquad+ADSR), {dt:.1f} s, "
                      f"{cores} threads; C port of the reference's x16 path (Rust toolchain absent), gcc {_ORACLE_FLAGS}"}


# stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner to stdout at
# NCCL_DEBUG=VERSION), so file descriptor 1 points at stderr for the whole run and the result line goes to a
# saved copy of the original stdout.
_RESULT_FD = None


def capture_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the C port: the Rust crate
    cannot be built in this image) on all host cores; each step is a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from synth2_b200 import bankgen
    cores = os.cpu_count() or 1
    nv = max(cores * 64, 1024)          # enough voices per thread that thread start-up does not show (same as cpu_baseline)
    voices = bankgen.make_bank(nv, RENDER_FRAMES, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    oracle = cpu_oracle()
    st = oracle.bank_init_states(voices)
    frames = step_frames(args.steps, args.block)
    for _ in range(args.warmup):
        oracle.bank_render(voices, st, SR, FILTER_BIQUAD, args.block, want_voices=False, want_bus=False, nthreads=cores)
    st = oracle.bank_init_states(voices)
    t0 = time.perf_counter()
    for fr in frames:
        oracle.bank_render(voices, st, SR, FILTER_BIQUAD, fr, want_voices=False, want_bus=False, nthreads=cores)
    dt = time.perf_counter() - t0
    value = nv * sum(frames) / dt
    sample = (f"{nv} voices x {sum(frames)} frames of the config-3 bank per run, {cores} threads, C port of the reference "
              f"x16 path, gcc {_ORACLE_FLAGS}")
    emit({
        "impl": "reference", "metric": "voice-samples/sec (osc+biquad, 48 kHz)", "value": value,
        "unit": "voice-samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 3: osc + resonant biquad + ADSR, 48 kHz, 4096-frame blocks (CPU sample of the 65,536-voice bank)",
                   "voices": nv, "block_frames": args.block},
        "cpu_baseline": {"value": value, "unit": "voice-samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voice-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------

def main():
    args = parse()
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import synth2_b200 as s2
    from synth2_b200 import bankgen

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    V, T = args.voices, args.block
    frames = step_frames(args.steps, T)
    total_frames = sum(frames)
    voices = bankgen.make_bank(V, RENDER_FRAMES, first_voice=rank * V, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    stream = torch.cuda.current_stream()
    bank = s2.VoiceBank(voices, SR, FILTER_BIQUAD, device=local, stream=stream)
    if args.pipeline > 1:
        bank.set_pipeline(args.pipeline)
    ring = [torch.empty((V, T), device=dev, dtype=torch.float32) for _ in range(2)]   # 2 x 1 GiB
    state0 = bank.get_state()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident render ------------------------------------------------------
    want_master = args.master_bus and world > 1
    master = torch.zeros(total_frames, device=dev, dtype=torch.float32) if want_master else None
    for i in range(max(args.warmup, 3)):
        bank.render(T, ring[i & 1], T, None)
    bank.set_state(state0)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = s2.lib().s2_launch_count()
    with ClockSampler(local) as clocks:
        ev0.record(stream)
        pos = 0
        for i, fr in enumerate(frames):
            bank.render(fr, ring[i & 1], T, master[pos:pos + fr] if want_master else None)
            pos += fr
        bank.join(stream)            # pipelined banks: the timing stream waits for every voice range
        if want_master:
            dist.reduce(master, dst=0, op=dist.ReduceOp.SUM)
        ev1.record(stream)
        barrier()
    launches = s2.lib().s2_launch_count() - launches0
    ms = rank_max(ev0.elapsed_time(ev1))
    value = world * V * total_frames / (ms * 1e-3)
    final_state = bank.get_state()
    assert np.all(final_state["frame_offset"] == total_frames)
    assert bool(torch.isfinite(ring[(len(frames) - 1) & 1][:, :frames[-1]]).all())

    # ---- e2e: host-buffer calls, H2D + D2H inside the timed region ----------------------------
    e2e = None
    if not args.no_e2e:
        rel_host = torch.from_numpy(voices["release_offset"].astype(np.uint32).view(np.int32)).pin_memory()
        bus_host = [torch.empty(T, dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        checksum = 0.0

        def e2e_pass(n_steps_frames):
            """Two buffers in flight, like the reference's player (audio_player.rs:56-60): enqueue step i,
            then wait for and consume step i-1."""
            nonlocal checksum
            for i, fr in enumerate(n_steps_frames):
                bank.set_releases(rel_host)                                              # H2D, pinned
                bank.render_bus_host_async(fr, bus_host[i & 1], ring[i & 1], T)          # render + mix + D2H
                done[i & 1].record(stream)
                if i > 0:
                    done[(i - 1) & 1].synchronize()
                    checksum += float(bus_host[(i - 1) & 1][0])                          # the host reads the result
            done[(len(n_steps_frames) - 1) & 1].synchronize()
            checksum += float(bus_host[(len(n_steps_frames) - 1) & 1][0])

        bank.set_state(state0)
        e2e_pass([T] * 3)
        bank.set_state(state0)
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        e2e_pass(frames)
        bank.join(stream)
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        assert np.isfinite(checksum)
        e_ms = rank_max(max(ev0.elapsed_time(ev1), wall * 1e3))
        e2e = {"value": world * V * total_frames / (e_ms * 1e-3), "unit": "voice-samples/s",
               "h2d_bytes_per_step": int(rel_host.numel() * 4), "d2h_bytes_per_step": int(T * 4),
               "ms_per_step": e_ms / len(frames),
               "what": "per step: note-off table H2D (pinned), render into the device ring + mono mix, mix D2H to "
                       "pinned host (the Synth::sample result), host waits for and reads the previous step's mix "
                       "(two buffers in flight, as the reference's audio_player does); wall clock"}

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    per_launch_bytes = V * T * BYTES_PER_VOICE_SAMPLE
    kernel_ms = ms / len(frames)            # one render kernel per step is the whole timed region
    achieved = V * total_frames * BYTES_PER_VOICE_SAMPLE / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "s2::render_kernel<1,1,0> (NV=1 voice/lane, FILTER=biquad)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": kernel_ms}
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            roofline["traffic"] = json.loads(prof.read_text()).get("render_kernel_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, bankgen)

    if rank == 0:
        emit({
            "metric": "voice-samples/sec (osc+biquad, 48 kHz)", "value": value, "unit": "voice-samples/s",
            "n_gpus": world, "steps": len(frames), "warmup": max(args.warmup, 3), "ms_per_step": ms / len(frames),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 3: 65,536 voices/GPU, osc (saw/square) + resonant biquad + ADSR, 48 kHz, "
                                   "60 s render in 4,096-frame blocks, state carried on device",
                       "voices_per_gpu": V, "block_frames": T, "render_frames": total_frames,
                       "l2_hygiene": "each step writes a fresh 1 GiB block (ring of 2) >> 126 MB L2; no input is re-read",
                       "master_bus": bool(want_master), "pipeline_voice_ranges": int(args.pipeline)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        })
    bank.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
