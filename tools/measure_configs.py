"""One-GPU measurements of the BASELINE configs that are not the bench line (bench.py is config 3).

  config 2   1,024 saw/square voices + one-pole low-pass, 4,096-frame buffers, 64+ consecutive blocks,
             16 rotating output buffers (a 16 MiB block would otherwise sit in the 126 MB L2)
  config 4   one GPU's 32,768-voice share of the 262,144-voice bank (8 GPUs), rows + master-bus mix
             (the NCCL reduce of the whole-render bus is measured by `bench.py --gpus N --master-bus`)
  config 5   one GPU's share of the sweep: 32,768 patch variants x 480,000 frames, every variant's
             f32 render kept on the device (62.9 GB of the 180 GB)

Prints one JSON line per config: voice-samples/s and the fraction of the measured HBM roofline at
4 algorithmic bytes per voice-sample.  `PYTHONPATH=. python tools/measure_configs.py [2] [2b] [4] [5]`; under
`torchrun --nproc-per-node 8` every rank measures its own GPU (config 5: the rank is the sweep's GPU index).
"""
import json
import os
import pathlib
import sys

import torch

import synth2_b200 as s2
from synth2_b200 import bankgen

ROOT = pathlib.Path(__file__).resolve().parent.parent
SR = 48000
# under torchrun every rank measures its own GPU (config 5: rank = the sweep's GPU index; no collective)
RANK = int(os.environ.get("RANK", "0"))
DEV = int(os.environ.get("LOCAL_RANK", "0"))


def peak_gbs():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(stream, bank, body):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    body()
    bank.join(stream)
    ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3


def report(name, voices, frames, seconds, **extra):
    vs = voices * frames / seconds
    # one write per line: the ranks of a torchrun share this stdout
    sys.stdout.write(json.dumps({"rank": RANK, "config": name, "voices": voices, "frames_per_voice": frames, "seconds": seconds,
                      "voice_samples_per_s": vs, "hbm_GBs": vs * 4 / 1e9, "frac_of_measured_hbm": vs * 4 / 1e9 / peak_gbs(),
                      **extra}) + "\n")
    sys.stdout.flush()


def config2(stream):
    V, T, blocks = 1024, 4096, 256
    voices = bankgen.make_bank(V, 4 * blocks * T)
    for time_split in (False, True):
        bank = s2.VoiceBank(voices, SR, s2.FILTER_ONE_POLE, device=DEV, stream=stream)
        if time_split:
            bank.set_time_split(True)
        ring = [torch.empty((V, T), device="cuda", dtype=torch.float32) for _ in range(16)]
        st = bank.get_state()
        for i in range(8):
            bank.render(T, ring[i & 15], T, None)
        bank.set_state(st)
        n0 = bank.time_split_blocks
        sec = timed(stream, bank, lambda: [bank.render(T, ring[i & 15], T, None) for i in range(blocks)])
        name = "2: 1,024 saw/square + one-pole, 4,096-frame buffers" + (", time-split" if time_split else ", one voice per lane")
        report(name + ", from note-on", V, blocks * T, sec, blocks=blocks, us_per_block=sec / blocks * 1e6,
               time_split_blocks=bank.time_split_blocks - n0)
        n0 = bank.time_split_blocks
        sec = timed(stream, bank, lambda: [bank.render(T, ring[i & 15], T, None) for i in range(blocks)])
        report(name + ", next 256 blocks (sustain)", V, blocks * T, sec, blocks=blocks, us_per_block=sec / blocks * 1e6,
               time_split_blocks=bank.time_split_blocks - n0)
        bank.close()


def config2_biquad(stream):
    """Config 2's shape with the resonant low-pass: the 2x2 scan of the time-split kernels."""
    V, T, blocks = 1024, 4096, 256
    voices = bankgen.make_bank(V, 4 * blocks * T, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    for time_split in (False, True):
        bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=DEV, stream=stream)
        if time_split:
            bank.set_time_split(True)
        ring = [torch.empty((V, T), device="cuda", dtype=torch.float32) for _ in range(16)]
        for i in range(16):
            bank.render(T, ring[i & 15], T, None)
        n0 = bank.time_split_blocks
        sec = timed(stream, bank, lambda: [bank.render(T, ring[i & 15], T, None) for i in range(blocks)])
        report("2b: 1,024 saw/square + biquad, 4,096-frame buffers, sustain" + (", time-split" if time_split else ", one voice per lane"),
               V, blocks * T, sec, blocks=blocks, us_per_block=sec / blocks * 1e6, time_split_blocks=bank.time_split_blocks - n0)
        bank.close()


def config4(stream):
    V, T, blocks = 32768, 4096, 118            # 10 s
    voices = bankgen.make_bank(V, blocks * T, mod_to_lpf_choices=bankgen.MOD_TO_LPF_BIQUAD)
    bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=DEV, stream=stream)
    bank.set_pipeline(4)
    ring = [torch.empty((V, T), device="cuda", dtype=torch.float32) for _ in range(2)]
    master = torch.empty(blocks * T, device="cuda", dtype=torch.float32)
    st = bank.get_state()
    for i in range(4):
        bank.render(T, ring[i & 1], T, master[:T])
    bank.set_state(st)
    sec = timed(stream, bank, lambda: [bank.render(T, ring[i & 1], T, master[i * T:(i + 1) * T]) for i in range(blocks)])
    report("4: one GPU's 32,768-voice share, rows + master-bus mix, 10 s", V, blocks * T, sec)
    sec = timed(stream, bank, lambda: [bank.render(T, None, 0, master[i * T:(i + 1) * T]) for i in range(blocks)])
    report("4 (mix only, no per-voice rows: no HBM stream)", V, blocks * T, sec)
    bank.close()


def config5(stream):
    V, total, T = bankgen.SWEEP_VARIANTS, 480000, 4096
    voices = bankgen.make_sweep_bank(RANK, total)
    bank = s2.VoiceBank(voices, SR, s2.FILTER_BIQUAD_LP, device=DEV, stream=stream)
    bank.set_pipeline(4)
    out = torch.empty((V, total), device="cuda", dtype=torch.float32)          # 62.9 GB
    st = bank.get_state()
    for i in range(3):
        bank.render(T, out[:, i * T:], total, None)
    bank.set_state(st)

    def body():
        pos = 0
        while pos < total:
            fr = min(T, total - pos)
            bank.render(fr, out[:, pos:], total, None)
            pos += fr
    sec = timed(stream, bank, body)
    # per-variant RMS as the summary a sweep reads back (outside the timed region; chunked to bound temporaries)
    rms = torch.empty(V, device="cuda", dtype=torch.float32)
    for a in range(0, V, 1024):
        rms[a:a + 1024] = out[a:a + 1024].double().pow(2).mean(dim=1).sqrt().float()
    finite = bool(torch.isfinite(rms).all())
    report("5: 32,768 patch variants x 10 s (one GPU's share), renders kept on device", V, total, sec,
           output_GB=V * total * 4 / 1e9, all_finite=finite, rms_min=float(rms.min()), rms_max=float(rms.max()))
    bank.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "2b", "4", "5"]
    torch.cuda.set_device(DEV)
    stream = torch.cuda.current_stream()
    for w in which:
        {"2": config2, "2b": config2_biquad, "4": config4, "5": config5}[w](stream)
