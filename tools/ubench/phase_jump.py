#!/usr/bin/env python3
"""Exact jump-ahead of the binary32 phase recurrence (try3/oscillators.rs:377-381), and why it does not lift
BASELINE config 2.

    phase' = (phase + d) % 1.0 in binary32, d = RN(1 / period)

While phase and phase + d lie in the same binade [2^k, 2^(k+1)) the rounded sum is phase + c_k ulp_k with the SAME
integer c_k = RN(d / ulp_k) every step (phase is a multiple of ulp_k; a tie can only occur when d's bits below
ulp_k are exactly 100...0, and then it alternates with the parity of the phase: those d step one at a time).  So m
steps inside a binade are one integer multiply-add on the mantissa.  Steps that change binade (or wrap) are taken
one at a time.  This script
  1. checks the jump against the sequential recurrence bit for bit over 4,096 frames for every note of the config-2
     bank (MIDI 24..108 at 48 kHz), and
  2. counts the serial steps (single steps + jumps) a lane needs for a 4,096-frame block.
Result (profiles/r2_phase_jump.txt): the count falls from 4,096 to a few hundred for low notes, but a lane's serial
work is still ~10 operations per jump against 3 per plain step, and above ~C6 (period < 46 frames) there are too few
steps per binade for a jump to pay at all.  `ts_phase_kernel` runs one voice per lane and a warp ends with its slowest
lane, and the config-2 bank draws notes uniformly from 24..108: every warp holds voices with periods of 6..40 frames,
which step 4,096 times at 14.75 cycles.  The pre-pass therefore stays at ~31 us per block whatever the low notes do.
"""
import numpy as np

F = np.float32
ONE = F(1.0)


def step(p, d):
    t = F(p + d)
    return F(t - ONE) if t >= ONE else t


def sequential(p, d, n):
    for _ in range(n):
        p = step(p, d)
    return p


def jump(p, d, n):
    """n steps from phase p; returns (phase, serial operations used)."""
    ops = 0
    while n > 0:
        if p > 0:
            k = int(np.floor(np.log2(float(p))))
            ulp = 2.0 ** (k - 23)
            hi = 2.0 ** (k + 1)
            ratio = float(d) / ulp
            c = np.floor(ratio + 0.5)
            tie = (ratio + 0.5) == c and (ratio != np.floor(ratio))
            if not tie and c >= 1 and float(p) + float(d) < hi:
                # steps that stay inside the binade: p + m * c * ulp + d < hi for the last of them
                m = int((hi - float(d) - float(p)) // (c * ulp)) + 1
                m = min(m, n)
                # the m-th result must still be below hi (it is: the (m-1)-th sum was)
                q = float(p) + m * c * ulp
                if q < hi and q < 1.0 and m >= 2:
                    p = F(q)
                    n -= m
                    ops += 1
                    continue
        p = step(p, d)
        n -= 1
        ops += 1
    return p, ops


def main():
    sr = 48000.0
    print("note  period   serial ops per 4096 frames   exact")
    worst = 0
    for note in range(24, 109, 6):
        pitch = F(440.0 * 2.0 ** ((note - 69) / 12.0))
        P = F(F(sr) / pitch)
        d = F(ONE / P)
        p0 = F(0.0)
        want = sequential(p0, d, 4096)
        got, ops = jump(p0, d, 4096)
        ok = np.float32(want).tobytes() == np.float32(got).tobytes()
        # also from a phase in the middle of a render
        p1 = sequential(p0, d, 12345)
        ok &= np.float32(sequential(p1, d, 4096)).tobytes() == np.float32(jump(p1, d, 4096)[0]).tobytes()
        worst = max(worst, ops)
        print(f"{note:4d} {float(P):8.2f} {ops:10d}                  {ok}")
        assert ok
    print(f"serial operations of the slowest lane: {worst} (plain stepping: 4096)")


if __name__ == "__main__":
    main()
