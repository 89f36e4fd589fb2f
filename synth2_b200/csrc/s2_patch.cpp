// s2_patch.cpp — the `.synth2` patch text format (host only, no CUDA).
//
// The reference ships `example.synth2` = `synth mySynth { }` and no reader for it: `Synth.config` is private
// and only ever `default_config()` (s2_lib/src/try3/synth.rs:10,56,125-152).  This reader gives the file
// the smallest body that can describe a patch: the fields of `static_config::Layer`
// (s2_lib/src/try3/static_config.rs:3-44) under their own names, everything optional, so the reference's
// example (an empty block) is the default patch.  A second, optional `score { }` block lists note events.
//
//   file   := { "synth" NAME "{" { field } "}" | "score" "{" { event } "}" }
//   field  := NAME ( "{" { field } "}" | VALUE ) [ ";" ]
//   event  := ( "on" TIME NOTE [ VELOCITY ] | "off" TIME NOTE ) [ ";" ]
//   TIME   := frames | <number>s | <number>ms          NOTE := 0..127
//   '#' and '//' start a comment.
//
//   synth lead {
//       osc { kind saw; gain 1.0 }              # kind: square | saw | triangle | sine
//       noise 0.0
//       lpf { freq 200; kind one_pole }         # kind: one_pole (filters.rs) | biquad | biquad_hp | biquad_bp |
//                                               #       first_order | first_order_hp (dsp_filters.rs); damping 1.414
//       amp_env { attack 100; decay 100; sustain 0.5; release 100 }          # Ms, Ms, Unipolar<1>, Ms
//       mod_env { attack 0; decay 200; sustain 0; release 0 }
//       modulations { mod_env_to_osc_freq 0; mod_env_to_lpf_freq 10 }        # Bipolar<10>
//   }
//
// Ranges follow units.rs:55-65 (`Unipolar<N>`: 0..=N, `Bipolar<N>`: -N..=N).
#include "../../include/s2_cuda.h"

#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace s2 {
int set_error(int code, const char* fmt, ...);       // s2_capi.cu: fills s2_last_error()
}

namespace {

struct Tok {
    enum Kind { End, Name, Number, LBrace, RBrace, Semi } kind = End;
    std::string text;
    double number = 0.0;
    char unit = 0;        // 0 = none, 's' = seconds, 'm' = milliseconds
    int line = 1;
};

struct Lexer {
    const char* p;
    int line = 1;
    explicit Lexer(const char* text) : p(text) {}

    void skip() {
        for (;;) {
            while (*p && isspace((unsigned char)*p)) { if (*p == '\n') line++; p++; }
            if (*p == '#' || (p[0] == '/' && p[1] == '/')) { while (*p && *p != '\n') p++; continue; }
            return;
        }
    }

    bool next(Tok& t, std::string& err) {
        skip();
        t = Tok();
        t.line = line;
        if (!*p) { t.kind = Tok::End; return true; }
        const char c = *p;
        if (c == '{') { t.kind = Tok::LBrace; p++; return true; }
        if (c == '}') { t.kind = Tok::RBrace; p++; return true; }
        if (c == ';') { t.kind = Tok::Semi; p++; return true; }
        if (isalpha((unsigned char)c) || c == '_') {
            const char* b = p;
            while (isalnum((unsigned char)*p) || *p == '_') p++;
            t.kind = Tok::Name;
            t.text.assign(b, p);
            return true;
        }
        if (isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
            char* end = nullptr;
            t.number = strtod(p, &end);
            if (end == p) { err = "malformed number"; return false; }
            t.text.assign(p, (size_t)(end - p));
            p = end;
            if (p[0] == 'm' && p[1] == 's') { t.unit = 'm'; p += 2; }
            else if (p[0] == 's' && !isalnum((unsigned char)p[1])) { t.unit = 's'; p += 1; }
            if (isalnum((unsigned char)*p) || *p == '_') { err = "unknown unit after number"; return false; }
            t.kind = Tok::Number;
            return true;
        }
        err = std::string("unexpected character '") + c + "'";
        return false;
    }
};

struct Parser {
    Lexer lx;
    Tok t;
    uint32_t sample_rate;
    s2_patch* out;
    s2_note_event* events;
    size_t events_cap, n_events = 0;
    bool seen_synth = false;
    int rc = S2_OK;

    Parser(const char* text, uint32_t sr, s2_patch* o, s2_note_event* ev, size_t cap)
        : lx(text), sample_rate(sr), out(o), events(ev), events_cap(cap) {}

    bool err(const char* fmt, ...) {
        char msg[384];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(msg, sizeof msg, fmt, ap);
        va_end(ap);
        rc = s2::set_error(S2_ERR_INVALID, "patch line %d: %s", t.line, msg);
        return false;
    }
    bool advance() {
        std::string e;
        if (!lx.next(t, e)) { t.line = lx.line; return err("%s", e.c_str()); }
        return true;
    }
    bool expect(Tok::Kind k, const char* what) {
        if (t.kind != k) return err("expected %s", what);
        return advance();
    }
    void opt_semi() { if (rc == S2_OK && t.kind == Tok::Semi) advance(); }

    bool number(const char* field, double lo, double hi, float* dst) {
        if (t.kind != Tok::Number || t.unit) return err("%s: expected a plain number", field);
        if (!(t.number >= lo && t.number <= hi)) return err("%s: %s is outside [%g, %g]", field, t.text.c_str(), lo, hi);
        *dst = (float)t.number;
        return advance();
    }

    bool adsr(const char* name, float* a, float* d, float* s, float* r) {
        if (!expect(Tok::LBrace, "'{'")) return false;
        while (t.kind == Tok::Name) {
            const std::string f = t.text;
            const std::string q = std::string(name) + "." + f;
            if (!advance()) return false;
            bool ok;
            if (f == "attack") ok = number(q.c_str(), 0.0, 1e9, a);
            else if (f == "decay") ok = number(q.c_str(), 0.0, 1e9, d);
            else if (f == "sustain") ok = number(q.c_str(), 0.0, 1.0, s);          // Unipolar<1>
            else if (f == "release") ok = number(q.c_str(), 0.0, 1e9, r);
            else return err("unknown field %s", q.c_str());
            if (!ok) return false;
            opt_semi();
            if (rc) return false;
        }
        return expect(Tok::RBrace, "'}'");
    }

    bool synth_block() {
        if (seen_synth) return err("more than one synth block");
        seen_synth = true;
        if (t.kind != Tok::Name) return err("expected the synth's name");
        snprintf(out->name, sizeof out->name, "%s", t.text.c_str());
        if (!advance() || !expect(Tok::LBrace, "'{'")) return false;
        s2_voice_desc& v = out->voice;
        while (t.kind == Tok::Name) {
            const std::string f = t.text;
            if (!advance()) return false;
            if (f == "osc") {
                if (!expect(Tok::LBrace, "'{'")) return false;
                while (t.kind == Tok::Name) {
                    const std::string g = t.text;
                    if (!advance()) return false;
                    if (g == "kind") {
                        if (t.kind != Tok::Name) return err("osc.kind: expected square, saw, triangle or sine");
                        if (t.text == "square") v.osc_kind = S2_OSC_SQUARE;
                        else if (t.text == "saw") v.osc_kind = S2_OSC_SAW;
                        else if (t.text == "triangle") v.osc_kind = S2_OSC_TRIANGLE;
                        else if (t.text == "sine") v.osc_kind = S2_OSC_SINE;
                        else return err("osc.kind: unknown oscillator '%s'", t.text.c_str());
                        if (!advance()) return false;
                    } else if (g == "gain") {
                        if (!number("osc.gain", 0.0, 1.0, &v.osc_gain)) return false;           // Unipolar<1>
                    } else return err("unknown field osc.%s", g.c_str());
                    opt_semi();
                    if (rc) return false;
                }
                if (!expect(Tok::RBrace, "'}'")) return false;
            } else if (f == "noise") {
                if (!number("noise", 0.0, 1.0, &v.noise_amt)) return false;                     // Unipolar<1>
            } else if (f == "lpf") {
                if (!expect(Tok::LBrace, "'{'")) return false;
                while (t.kind == Tok::Name) {
                    const std::string g = t.text;
                    if (!advance()) return false;
                    if (g == "freq") {
                        if (!number("lpf.freq", 0.0, 1e6, &v.lpf_freq_hz)) return false;
                    } else if (g == "damping") {
                        if (!number("lpf.damping", 0.0, 10.0, &v.damping)) return false;        // Unipolar<10>, dsp_filters.rs:95 (> 0 for the second-order kinds: s2_synth_set_patch)
                    } else if (g == "kind") {
                        if (t.kind != Tok::Name) return err("lpf.kind: expected one_pole or biquad");
                        if (t.text == "one_pole") out->filter_kind = S2_FILTER_ONE_POLE;
                        else if (t.text == "biquad") out->filter_kind = S2_FILTER_BIQUAD_LP;
                        else if (t.text == "biquad_hp") out->filter_kind = S2_FILTER_BIQUAD_HP;
                        else if (t.text == "biquad_bp") out->filter_kind = S2_FILTER_BIQUAD_BP;       // damping = quality factor
                        else if (t.text == "first_order") out->filter_kind = S2_FILTER_FIRST_ORDER_LP;
                        else if (t.text == "first_order_hp") out->filter_kind = S2_FILTER_FIRST_ORDER_HP;
                        else return err("lpf.kind: unknown filter '%s'", t.text.c_str());
                        if (!advance()) return false;
                    } else return err("unknown field lpf.%s", g.c_str());
                    opt_semi();
                    if (rc) return false;
                }
                if (!expect(Tok::RBrace, "'}'")) return false;
            } else if (f == "amp_env") {
                if (!adsr("amp_env", &v.amp_attack_ms, &v.amp_decay_ms, &v.amp_sustain, &v.amp_release_ms)) return false;
            } else if (f == "mod_env") {
                if (!adsr("mod_env", &v.mod_attack_ms, &v.mod_decay_ms, &v.mod_sustain, &v.mod_release_ms)) return false;
            } else if (f == "modulations") {
                if (!expect(Tok::LBrace, "'{'")) return false;
                while (t.kind == Tok::Name) {
                    const std::string g = t.text;
                    if (!advance()) return false;
                    bool ok;
                    if (g == "mod_env_to_osc_freq") ok = number("modulations.mod_env_to_osc_freq", -10.0, 10.0, &v.mod_env_to_osc_freq);
                    else if (g == "mod_env_to_lpf_freq") ok = number("modulations.mod_env_to_lpf_freq", -10.0, 10.0, &v.mod_env_to_lpf_freq);
                    else return err("unknown field modulations.%s", g.c_str());
                    if (!ok) return false;
                    opt_semi();
                    if (rc) return false;
                }
                if (!expect(Tok::RBrace, "'}'")) return false;
            } else {
                return err("unknown field %s", f.c_str());
            }
            opt_semi();
            if (rc) return false;
        }
        return expect(Tok::RBrace, "'}'");
    }

    bool time_value(uint64_t* frame) {
        if (t.kind != Tok::Number) return err("expected a time (frames, or a number followed by s / ms)");
        if (!(t.number >= 0.0)) return err("negative time");
        if (t.unit && sample_rate == 0) return err("times in s / ms need a sample rate");
        double f = t.number;
        if (t.unit == 's') f = t.number * (double)sample_rate;
        else if (t.unit == 'm') f = t.number * (double)sample_rate / 1000.0;
        else if (f != floor(f)) return err("a time without a unit is a whole number of frames");
        if (f > 1.8e19) return err("time out of range");
        *frame = (uint64_t)llround(f);
        return advance();
    }

    bool score_block() {
        if (!expect(Tok::LBrace, "'{'")) return false;
        uint64_t last = 0;
        while (t.kind == Tok::Name) {
            const bool on = t.text == "on";
            if (!on && t.text != "off") return err("expected 'on' or 'off'");
            if (!advance()) return false;
            s2_note_event ev;
            memset(&ev, 0, sizeof ev);
            ev.on = on ? 1 : 0;
            ev.velocity = 1.0f;
            if (!time_value(&ev.frame)) return false;
            if (ev.frame < last) return err("events must be in time order");
            last = ev.frame;
            if (t.kind != Tok::Number || t.unit || t.number != floor(t.number) || t.number < 0 || t.number > 127)
                return err("expected a MIDI note number 0..127");
            ev.note = (uint8_t)t.number;
            if (!advance()) return false;
            if (on && t.kind == Tok::Number) {
                if (!number("velocity", 0.0, 1.0, &ev.velocity)) return false;               // Velocity(Unipolar<1>)
            }
            if (events) {
                if (n_events >= events_cap) return err("more than %zu events", events_cap);
                events[n_events] = ev;
            }
            n_events++;
            opt_semi();
            if (rc) return false;
        }
        return expect(Tok::RBrace, "'}'");
    }

    int run() {
        if (!advance()) return rc;
        while (t.kind != Tok::End) {
            if (t.kind != Tok::Name) { err("expected 'synth' or 'score'"); return rc; }
            const std::string kw = t.text;
            if (!advance()) return rc;
            if (kw == "synth") { if (!synth_block()) return rc; }
            else if (kw == "score") { if (!score_block()) return rc; }
            else { err("expected 'synth' or 'score', found '%s'", kw.c_str()); return rc; }
        }
        if (!seen_synth) { t.line = lx.line; err("no synth block"); return rc; }
        return S2_OK;
    }
};

}  // namespace

extern "C" {

void s2_default_patch(s2_patch* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    s2_default_voice(&p->voice);
    p->filter_kind = S2_FILTER_ONE_POLE;
}

int s2_patch_parse(const char* text, uint32_t sample_rate, s2_patch* out, s2_note_event* events, size_t events_cap,
                   size_t* n_events) {
    if (n_events) *n_events = 0;
    if (!text || !out) return s2::set_error(S2_ERR_INVALID, "null argument");
    s2_default_patch(out);
    Parser ps(text, sample_rate, out, events, events_cap);
    const int rc = ps.run();
    if (rc) return rc;
    if (n_events) *n_events = ps.n_events;
    return S2_OK;
}

}  // extern "C"
