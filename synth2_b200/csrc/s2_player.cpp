// s2_player.cpp — the hand-off between the renderer and an audio callback (host only).
//
// Mirrors s2_bin/src/audio_player.rs + the synth thread of main.rs without the cpal device:
//   * two mono buffers of BUFFER_FRAMES = 2048 frames circulate (audio_player.rs:21-24, 56-60): the synth
//     thread takes an empty one, applies the note messages that arrived, renders, and hands it over filled
//     (main.rs:134-160); here the render is one GPU launch + one copy into pinned host memory;
//   * the audio callback (`fill_buffer`, audio_player.rs:136-199) first drains the buffer it already holds,
//     then takes AT MOST ONE new filled buffer without blocking, writes every sample to all output channels,
//     and pads what is left with zeros (an underrun is logged, never waited for); a drained buffer goes back
//     to the synth thread (audio_player.rs:205-255).
// The audio side takes no lock, as the reference's does not (`try_recv` / `send` on channels of depth 2): buffers
// travel through two single-producer / single-consumer rings of atomics, the synth thread's error state is an atomic
// flag, and the only thing the callback does beyond loads and stores is a condition-variable notify (no mutex held).
// The synth thread sleeps on that condition variable with a 1 ms timeout, so a notify that races its check costs at
// most a millisecond of a 42.7 ms buffer.
// The reference polls MIDI between 16-frame chunks while it fills a buffer (main.rs:138-147); a buffer
// renders here in microseconds, so messages are applied once, before the buffer is rendered.
#include "../../include/s2_cuda.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

namespace s2 {
int set_error(int code, const char* fmt, ...);
}

namespace {
constexpr size_t kFrames = S2_PLAYER_BUFFER_FRAMES;
constexpr int kBuffers = 2;            // sync_channel(2) each way, two buffers in circulation

struct Msg { uint8_t note; uint8_t on; float velocity; };

// Single-producer / single-consumer ring of buffer indices (capacity 4 >= the two buffers in circulation).
struct Ring {
    std::atomic<uint32_t> head{0}, tail{0};      // pop at head, push at tail
    int slot[4] = {};
    bool push(int v) {                           // producer only
        const uint32_t t = tail.load(std::memory_order_relaxed);
        if (t - head.load(std::memory_order_acquire) == 4u) return false;
        slot[t & 3u] = v;
        tail.store(t + 1u, std::memory_order_release);
        return true;
    }
    bool pop(int* v) {                           // consumer only
        const uint32_t h = head.load(std::memory_order_relaxed);
        if (h == tail.load(std::memory_order_acquire)) return false;
        *v = slot[h & 3u];
        head.store(h + 1u, std::memory_order_release);
        return true;
    }
    bool empty() const { return head.load(std::memory_order_acquire) == tail.load(std::memory_order_acquire); }
};
}  // namespace

struct s2_player {
    int device = 0;
    uint32_t sample_rate = 0;
    s2_synth* synth = nullptr;
    float* buf[kBuffers] = {};         // pinned host memory
    // synth thread <-> audio callback: indices of buffers, FIFO each way, lock-free
    Ring empty_q;                      // callback -> synth thread
    Ring filled_q;                     // synth thread -> callback
    std::mutex mu;                     // control side only (note messages, start / stop, waiters): never the callback
    std::condition_variable cv_empty, cv_progress;
    std::vector<Msg> inbox;            // note messages not yet applied
    bool started = false, stop = false;
    std::atomic<int> failed{0};        // error code of the synth thread, sticky; fail_msg is written before it is set
    std::string fail_msg;
    std::atomic<uint64_t> rendered{0}; // buffers handed over filled
    std::thread worker;
    // consumer-only state (the callback thread)
    int pending = -1;                  // buffer being drained
    size_t consumed = 0;
    std::atomic<uint64_t> underruns{0}, frames_played{0};
};

namespace {

void synth_thread(s2_player* p) {
    for (;;) {
        int b;
        std::vector<Msg> msgs;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            // the callback pushes into empty_q and notifies WITHOUT this mutex: poll at 1 ms so a notify that
            // slips between the check and the wait is not waited for
            while (!(p->stop || (p->started && !p->empty_q.empty())))
                p->cv_empty.wait_for(lk, std::chrono::milliseconds(1));
            if (p->stop) return;
            p->empty_q.pop(&b);
            msgs.swap(p->inbox);
        }
        int rc = S2_OK;
        for (const Msg& m : msgs) {
            rc = m.on ? s2_synth_note_on(p->synth, m.note, m.velocity) : s2_synth_note_off(p->synth, m.note);
            if (rc < 0) break;
        }
        if (rc >= 0) rc = s2_synth_sample(p->synth, p->buf[b], kFrames, p->sample_rate);
        if (rc < 0) {
            {
                std::lock_guard<std::mutex> lk(p->mu);
                p->fail_msg = s2_last_error();
                p->failed.store(rc, std::memory_order_release);
            }
            p->cv_progress.notify_all();
            return;
        }
        p->filled_q.push(b);
        {
            std::lock_guard<std::mutex> lk(p->mu);       // pairs with the waiters of s2_player_wait_buffers
            p->rendered.fetch_add(1, std::memory_order_release);
        }
        p->cv_progress.notify_all();
    }
}

// audio_player.rs:205-255: copy from the held buffer, mono -> every channel; hand it back when drained
size_t fill_from_pending(s2_player* p, float* out, size_t frames, uint32_t channels) {
    if (p->pending < 0) return 0;
    const size_t n = frames < kFrames - p->consumed ? frames : kFrames - p->consumed;
    const float* src = p->buf[p->pending] + p->consumed;
    for (size_t i = 0; i < n; i++)
        for (uint32_t c = 0; c < channels; c++) out[i * channels + c] = src[i];
    p->consumed += n;
    if (p->consumed == kFrames) {
        p->empty_q.push(p->pending);               // lock-free; two buffers never overfill a ring of four
        p->pending = -1;
        p->consumed = 0;
        p->cv_empty.notify_one();                  // no mutex: see synth_thread
    }
    return n;
}

}  // namespace

extern "C" {

int s2_player_new(int device, uint32_t sample_rate, s2_player** out) {
    if (!out) return s2::set_error(S2_ERR_INVALID, "null out");
    *out = nullptr;
    if (sample_rate == 0) return s2::set_error(S2_ERR_INVALID, "sample_rate must be > 0");
    s2_player* p = new (std::nothrow) s2_player;
    if (!p) return s2::set_error(S2_ERR_NOMEM, "out of host memory");
    p->device = device;
    p->sample_rate = sample_rate;
    int rc = s2_synth_new(device, &p->synth);
    if (rc) { delete p; return rc; }
    cudaSetDevice(device);
    for (int i = 0; i < kBuffers; i++) {
        if (cudaMallocHost(&p->buf[i], kFrames * sizeof(float)) != cudaSuccess) {
            for (int j = 0; j < i; j++) cudaFreeHost(p->buf[j]);
            s2_synth_free(p->synth);
            delete p;
            return s2::set_error(S2_ERR_NOMEM, "cudaMallocHost failed for the player buffers");
        }
        memset(p->buf[i], 0, kFrames * sizeof(float));       // Buffer(Box::from([0_f32; BUFFER_FRAMES]))
        p->empty_q.push(i);
    }
    p->worker = std::thread(synth_thread, p);
    *out = p;
    return S2_OK;
}

void s2_player_free(s2_player* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_empty.notify_all();
    if (p->worker.joinable()) p->worker.join();
    s2_synth_free(p->synth);
    cudaSetDevice(p->device);
    for (int i = 0; i < kBuffers; i++) cudaFreeHost(p->buf[i]);
    delete p;
}

int s2_player_set_patch(s2_player* p, const s2_patch* patch) {
    if (!p || !patch) return s2::set_error(S2_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(p->mu);
    if (p->started) return s2::set_error(S2_ERR_INVALID, "set the patch before s2_player_start");
    return s2_synth_set_patch(p->synth, patch);
}

int s2_player_start(s2_player* p) {
    if (!p) return s2::set_error(S2_ERR_INVALID, "null player");
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->started = true;
    }
    p->cv_empty.notify_all();
    return S2_OK;
}

static int post(s2_player* p, uint8_t note, uint8_t on, float velocity) {
    if (!p) return s2::set_error(S2_ERR_INVALID, "null player");
    std::lock_guard<std::mutex> lk(p->mu);
    if (const int f = p->failed.load(std::memory_order_acquire)) return s2::set_error(f, "synth thread: %s", p->fail_msg.c_str());
    p->inbox.push_back({note, on, velocity});
    return S2_OK;
}

int s2_player_note_on(s2_player* p, uint8_t note, float velocity) { return post(p, note, 1, velocity); }
int s2_player_note_off(s2_player* p, uint8_t note) { return post(p, note, 0, 0.0f); }

// audio_player.rs:136-199 `fill_buffer` for f32 output.  Never blocks on the renderer.
int64_t s2_player_fill(s2_player* p, float* out, size_t frames, uint32_t channels) {
    if (!p || (!out && frames)) return s2::set_error(S2_ERR_INVALID, "null argument");
    if (channels == 0) return s2::set_error(S2_ERR_INVALID, "channels must be > 0");
    size_t written = fill_from_pending(p, out, frames, channels);
    if (written < frames) {
        int got = -1;
        if (!p->filled_q.pop(&got)) got = -1;            // try_recv: lock-free
        const int failed = p->failed.load(std::memory_order_acquire);
        if (got >= 0) {
            p->pending = got;
            p->consumed = 0;
            written += fill_from_pending(p, out + written * channels, frames - written, channels);
        } else {
            p->underruns.fetch_add(1, std::memory_order_relaxed);   // "didn't receive buffer in time for audio out"
        }
        memset(out + written * channels, 0, (frames - written) * channels * sizeof(float));
        if (got < 0 && failed) return s2::set_error(failed, "synth thread: %s", p->fail_msg.c_str());   // written before `failed`
    }
    p->frames_played.fetch_add(written, std::memory_order_relaxed);
    return (int64_t)written;
}

int s2_player_stats(s2_player* p, uint64_t* buffers_rendered, uint64_t* underruns, uint64_t* frames_played) {
    if (!p) return s2::set_error(S2_ERR_INVALID, "null player");
    if (buffers_rendered) *buffers_rendered = p->rendered.load(std::memory_order_acquire);
    if (underruns) *underruns = p->underruns.load(std::memory_order_relaxed);
    if (frames_played) *frames_played = p->frames_played.load(std::memory_order_relaxed);
    return S2_OK;
}

int s2_player_wait_buffers(s2_player* p, uint64_t n_rendered, uint32_t timeout_ms) {
    if (!p) return s2::set_error(S2_ERR_INVALID, "null player");
    std::unique_lock<std::mutex> lk(p->mu);
    const bool ok = p->cv_progress.wait_for(lk, std::chrono::milliseconds(timeout_ms),
                                            [&] { return p->failed.load(std::memory_order_acquire) || p->rendered.load(std::memory_order_acquire) >= n_rendered; });
    if (const int f = p->failed.load(std::memory_order_acquire)) return s2::set_error(f, "synth thread: %s", p->fail_msg.c_str());
    return ok ? S2_OK : 1;
}

}  // extern "C"
