// render_patch.cpp — the C ABI of include/s2_cuda.h from compiled code, no Python anywhere:
// read a .synth2 patch (with an optional score), render it on the GPU, write raw little-endian f32.
//
//   g++ -std=c++17 -O2 -Iinclude examples/render_patch.cpp -Lsynth2_b200 -ls2cuda -Wl,-rpath,$PWD/synth2_b200 -o render_patch
//   ./render_patch example.synth2 10 48000 out.f32
//
// This is what `s2_bin render <patch> --seconds 10 --rate 48000 -o out.f32` would be (BASELINE config 1 names
// that entry; the reference's s2_bin has only `midi` and `build-tables`, s2_bin/src/main.rs:15-19).
#include "s2_cuda.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

static int fail(const char* what) {
    fprintf(stderr, "%s: %s\n", what, s2_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc != 5) {
        fprintf(stderr, "usage: %s PATCH.synth2 SECONDS RATE OUT.f32\n", argv[0]);
        return 2;
    }
    std::ifstream in(argv[1]);
    if (!in) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::stringstream text;
    text << in.rdbuf();
    const double seconds = atof(argv[2]);
    const uint32_t rate = (uint32_t)atoi(argv[3]);
    const size_t frames = (size_t)(seconds * rate + 0.5);

    s2_patch patch;
    size_t n_events = 0;
    if (s2_patch_parse(text.str().c_str(), rate, &patch, nullptr, 0, &n_events)) return fail("patch");
    std::vector<s2_note_event> events(n_events);
    if (n_events && s2_patch_parse(text.str().c_str(), rate, &patch, events.data(), n_events, &n_events)) return fail("patch");
    if (events.empty()) {                       // no score block: A4 held for three quarters of the render
        s2_note_event on = {}, off = {};
        on.note = off.note = 69; on.on = 1; on.velocity = 1.0f;
        off.frame = (uint64_t)(frames * 3 / 4);
        events = {on, off};
    }

    s2_synth* synth = nullptr;
    if (s2_synth_new(0, &synth)) return fail("s2_synth_new");
    if (s2_synth_set_patch(synth, &patch)) return fail("s2_synth_set_patch");
    std::vector<float> out(frames);
    if (s2_synth_render_score(synth, events.data(), events.size(), rate, out.data(), frames)) return fail("render");
    s2_synth_free(synth);

    FILE* f = fopen(argv[4], "wb");
    if (!f || fwrite(out.data(), sizeof(float), out.size(), f) != out.size()) { fprintf(stderr, "cannot write %s\n", argv[4]); return 2; }
    fclose(f);
    float peak = 0.0f;
    for (float v : out) peak = v > peak ? v : (-v > peak ? -v : peak);
    fprintf(stderr, "%s: %zu frames at %u Hz, %zu events, peak %.4f -> %s\n", patch.name, frames, rate, events.size(), peak, argv[4]);
    return 0;
}
