// check_cutoff.cpp — CPU check of synth2_b200/csrc/s2_cutoff.h (the scalar forms of what the kernels compile).
//   g++ -O2 -std=c++17 -ffp-contract=off -march=x86-64-v3 -o check_cutoff tools/check_cutoff.cpp && ./check_cutoff [quick]
// 1. div_in_range against the IEEE quotient over the operand range of the second-order filters, with the reciprocal
//    estimate off by up to +-2 ulp (the special-function unit's error is not modelled more finely than that): how
//    often it is not the correctly rounded value, and by how much;
// 2. exp_neg_fast against exp(-theta) in binary64;
// 3. windowed sin / cos against the once-rounded binary64 values: mismatch rates and worst absolute error;
// 4. a whole decay sweep of the resonant low-pass at its worst corner (100-200 Hz, damping 0.2): coefficients from
//    windows vs the full evaluation, and the filter outputs they produce.
// Exits non-zero if 1 is off by more than an ulp or wrong in more than 1e-4 of the cases, if 2 exceeds 4 ulp, if 3 adds
// more than 4e-9 to a correct rounding (or misses the rounded-once value in more than 2 % of slow-sweep frames), or if
// 4 changes a coefficient in more than 0.1 % of the frames.
#include "../synth2_b200/csrc/s2_cutoff.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

using namespace s2c;

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float fbits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void make_window(Window& W, float thc) {
    W.k = 0; W.valid = 1; W.thc = thc;
    double s, c; s2_sincos_d((double)thc, &s, &c);
    split_hi_lo(s, &W.Ah, &W.Al); split_hi_lo(c, &W.Bh, &W.Bl);
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    int fail = 0;

    // ---- 1. division
    {
        std::mt19937_64 rng(12345);
        long n = 0, bad = 0, worse = 0;
        const long N = quick ? 4000000 : 40000000;
        for (long i = 0; i < N; i++) {
            const float h = (float)((rng() >> 11) * (1.0 / 9007199254740992.0)) * 8.0f;      // hd * sin in [0, 8)
            const float num = 1.0f - h, den = 1.0f + h, nden = -1.0f - h;
            const float want = num / den;
            const float r0 = 1.0f / den;
            for (int p = -2; p <= 2; p++) {
                const float r = fbits(bits(r0) + (uint32_t)p);
                const float got = div_in_range_from<float>(num, nden, r);
                n++;
                if (bits(got) != bits(want)) {
                    bad++;
                    const int32_t du = (int32_t)(bits(got) - bits(want));
                    worse += du > 1 || du < -1;
                }
            }
        }
        printf("div_in_range: %ld cases (reciprocal estimate off by -2..+2 ulp), %ld not the IEEE quotient (%.2e), %ld off by more than an ulp\n",
               n, bad, (double)bad / n, worse);
        fail |= worse != 0 || bad * 10000 > n;
    }

    // ---- 2. one-pole k = e^-theta
    {
        std::mt19937_64 rng(99);
        double worst = 0.0;
        const long N = quick ? 400000 : 4000000;
        for (long i = 0; i < N; i++) {
            const double u = (rng() >> 11) * (1.0 / 9007199254740992.0);
            const float th = (float)(0.004 * exp(u * log(40.0 / 0.004)));
            const float k = exp_neg_fast<float>(th);
            const double want = exp(-(double)th);
            const double ulp = (double)(fbits(bits((float)want) + 1u) - (float)want);
            const double e = fabs((double)k - want) / ulp;
            if (e > worst) worst = e;
        }
        printf("exp_neg_fast: worst error %.2f ulp over theta in [0.004, 40] (host exp2f standing in for ex2.approx)\n", worst);
        fail |= worst > 4.0;
    }

    // ---- 3. windows.  "slack" = |result - exact| - ulp(result)/2: what the evaluation adds on top of a correct rounding.
    for (int pass = 0; pass < 2; pass++) {
        // pass 0: sweeps as slow as the bench bank's (|d| <= 2 % of theta); pass 1: anything a valid window admits
        std::mt19937_64 rng(777);
        long n = 0, bs = 0, bc = 0;
        double ws = 0, wc = 0;
        const long N = quick ? 200000 : 2000000;
        for (long i = 0; i < N; i++) {
            const double u = (rng() >> 11) * (1.0 / 9007199254740992.0);
            const float thc = (float)(0.004 * exp(u * log(3.1 / 0.004)));              // log-uniform 0.004 .. 3.1
            Window W;
            make_window(W, thc);
            for (int j = 0; j < 16; j++) {
                const double v = (rng() >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
                const float lim = thc * (pass == 0 ? 0.02f : 0.5f);
                const float dmax = lim < kWinDelta ? lim : kWinDelta;
                const float th = thc + (float)(v * dmax);
                float s, c;
                window_sincos<float>(W, th - thc, &s, &c);                 // exact difference (Sterbenz)
                const double sd = sin((double)th), cd = cos((double)th);
                n++;
                bs += bits(s) != bits((float)sd); bc += bits(c) != bits((float)cd);
                auto slack = [](float got, double want) {
                    const double half_ulp = 0.5 * (double)(fbits(bits(fabsf(got)) + 1u) - fabsf(got));
                    const double x = fabs((double)got - want) - half_ulp;
                    return x > 0.0 ? x : 0.0;
                };
                const double es = slack(s, sd), ec = slack(c, cd);
                if (es > ws) ws = es;
                if (ec > wc) wc = ec;
            }
        }
        printf("windows (%s): %ld frames; != rounded-once: sin %.3f%% cos %.3f%%; worst slack beyond a correct "
               "rounding: sin %.1e cos %.1e\n", pass == 0 ? "|d| <= 2% of theta" : "|d| <= 50% of theta, 2^-7",
               n, 100.0 * bs / n, 100.0 * bc / n, ws, wc);
        fail |= ws > 4e-9 || wc > 4e-9;
        if (pass == 0) fail |= bs > n / 50 || bc > n / 50;
    }

    // ---- 4. a decay sweep at the worst corner: cutoff 100..200 Hz, damping 0.2, 1.5 octaves over 9600 frames
    {
        const float sr = 48000.0f, one = 1.0f, hd = 0.1f;
        double worst_out = 0.0;
        long coef_diff = 0, frames = 0;
        for (int voice = 0; voice < (quick ? 8 : 64); voice++) {
            const float lpf = 100.0f + 1.7f * (float)voice;
            const float P = sr / (55.0f * (1.0f + 0.37f * (float)voice));
            const float A = 0.0f, D = 9600.0f, sD = (0.0f - 1.0f) / D;
            float ph = 0.0f;
            float xa1 = 0, xa2 = 0, ya1 = 0, ya2 = 0, xb1 = 0, xb2 = 0, yb1 = 0, yb2 = 0;
            Window W; W.k = 0xffffffffu;
            for (uint32_t n = 0; n < 9600; n++) {
                const float x = (float)n;
                const float m = (sD * (x - A)) + 1.0f;
                const float th0 = (kTwoPi * lpf) / sr;
                const float th = theta_at<float>(m, 1.5f, th0);
                if ((n >> 5) != W.k) {
                    const float xc = (float)((n & ~31u) + 16u);
                    const float thc = theta_at<float>((sD * (xc - A)) + 1.0f, 1.5f, th0);
                    make_window(W, thc); W.k = n >> 5;
                }
                float s, c, a0, a1, a2, b0, b1, b2;
                window_sincos<float>(W, delta_at<float>(m, 1.5f, th0, W.thc), &s, &c);
                biquad_lp_hp<false, float>(s, c, hd, one, &a0, &a1, &a2);
                // the full evaluation at the same angle: the window's delta is taken from the UNROUNDED product
                // sweep * theta0 (delta_at is one fma), so the comparison angle is that product in binary64
                const double thd = (double)sweep_at<float>(m, 1.5f) * (double)th0;
                (void)th;
                double sd, cd; s2_sincos_d(thd, &sd, &cd);
                biquad_lp_hp_any<false>((float)sd, (float)cd, hd, one, &b0, &b1, &b2);
                coef_diff += bits(a0) != bits(b0) || bits(a1) != bits(b1) || bits(a2) != bits(b2);
                frames++;
                // saw + 1 (the x16 path's input, process.rs:341-345), both filters (dsp_filters.rs:116-128)
                const float u = (1.0f - 2.0f * ph) + 1.0f;
                ph += 1.0f / P; if (ph >= 1.0f) ph -= 1.0f;
                float sx = (u + 2.0f * xa1) + xa2;
                float t = a0 * sx; t = t + a2 * ya1; t = t - a1 * ya2;
                xa2 = xa1; xa1 = u; ya2 = ya1; ya1 = t;
                sx = (u + 2.0f * xb1) + xb2;
                float w = b0 * sx; w = w + b2 * yb1; w = w - b1 * yb2;
                xb2 = xb1; xb1 = u; yb2 = yb1; yb1 = w;
                const double dv = fabs((double)t - (double)w);
                if (dv > worst_out) worst_out = dv;
            }
        }
        printf("decay sweep: %ld frames, %ld with a coefficient differing from the full evaluation (%.3f%%), worst output "
               "difference %.2e\n", frames, coef_diff, 100.0 * coef_diff / frames, worst_out);
        // (one differing ulp in one frame is enough to decorrelate the two filters' rounding, after which they sit
        // apart by the binary32 direct form's own noise floor at this corner, ~5e-5 rms: the output difference is
        // reported, the coefficient mismatch rate is what is asserted — glibc's sinf / cosf, which the oracle
        // calls, differ from the rounded-once values in 1.3 % of arguments)
        fail |= coef_diff * 1000 > frames;
    }
    // ---- 5. interpolation: (q, cos) at the frames of the grid of 4, linear in between, against the window evaluation of
    //         every frame — relative difference of c0 = 2 alpha (the coefficient that cancels) at the bench bank's sweep
    //         rate and at the rate limit kInterpRate12
    for (int pass = 0; pass < 2; pass++) {
        const float sr = 48000.0f, one = 1.0f;
        const float amt = 1.5f, D = pass == 0 ? 9600.0f : 1.5f * 12.0f / (kInterpRate12 * 1.0001f);      // 12 |amt es| at the limit
        const float sD = (0.0f - 1.0f) / D;
        double worst_rel = 0.0, worst_abs = 0.0, bias_sum = 0.0;
        long bias_n = 0;
        for (int voice = 0; voice < 24; voice++) {
            const float lpf = 100.0f * powf(80.0f, (float)voice / 23.0f), hd = 0.1f + 0.026f * (float)voice;
            const float th0 = (kTwoPi * lpf) / sr;
            const uint32_t frames = (uint32_t)D & ~31u;
            for (uint32_t w0 = 0; w0 + 32 <= frames; w0 += 32) {
                Window W;
                const float xc = (float)(w0 + 16u);
                make_window(W, theta_at<float>((sD * (xc - 0.0f)) + 1.0f, amt, th0));
                if (!(W.thc <= kThetaMax)) continue;
                for (uint32_t k = w0; k < w0 + 32; k += 4) {
                    float qa, ca, qb, cb;
                    node_q_cos<float>(W, (sD * (float)k) + 1.0f, amt, th0, hd, one, &qa, &ca);
                    node_q_cos<float>(W, (sD * (float)(k + 4)) + 1.0f, amt, th0, hd, one, &qb, &cb);
                    const float dq = (qb - qa) * 0.25f, dc = (cb - ca) * 0.25f;
                    for (uint32_t j = 0; j < 4; j++) {
                        float a0, a1, a2, b0, b1, b2, q, co;
                        biquad_from_q_cos<false, float>(fmaf((float)j, dq, qa), fmaf((float)j, dc, ca), one, &a0, &a1, &a2);
                        node_q_cos<float>(W, (sD * (float)(k + j)) + 1.0f, amt, th0, hd, one, &q, &co);
                        biquad_from_q_cos<false, float>(q, co, one, &b0, &b1, &b2);
                        const double rel = fabs((double)a0 - (double)b0) / fabs((double)b0);
                        if (rel > worst_rel) worst_rel = rel;
                        if (lpf < 250.0f) { bias_sum += ((double)a0 - (double)b0) / (double)b0; bias_n++; }
                        const double ab = fmax(fabs((double)a1 - (double)b1), fabs((double)a2 - (double)b2));
                        if (ab > worst_abs) worst_abs = ab;
                    }
                }
            }
        }
        // the per-frame evaluation itself scatters by ~2^-25 / alpha in c0 (the cancellation: ~7e-4 at 100 Hz); what the
        // interpolation may add is a BIAS of (4 ln2 |amt es|)^2 / 2 at most — the mean over the low cutoffs shows it
        const double bias = bias_n ? bias_sum / (double)bias_n : 0.0;
        printf("interpolation (%s): c0 relative difference worst %.2e (scatter of the cancellation), mean at cutoffs < 250 Hz %+.2e; "
               "c1 / c2 absolute difference worst %.2e\n", pass == 0 ? "1.5 octaves in 200 ms" : "rate limit", worst_rel, bias, worst_abs);
        fail |= worst_abs > (pass == 0 ? 4e-6 : 6e-5) || worst_rel > 2e-3 || fabs(bias) > (pass == 0 ? 2e-6 : 2e-5);
    }
    printf(fail ? "FAILED\n" : "ok\n");
    return fail;
}
