// The reference's producer/consumer FIFO (fifo.h:19-62, fifo.c) rebuilt behind the
// same names and semantics, so that gps_thread_ep's acquire/enqueue calls
// (gps.c:2698,2860-2865) and the sinks' dequeue/release calls (sdr_iqfile.c:38-49,
// sdr_hackrf.c:236-248, sdr_pluto.c:57-68) link against libgpsb200.so unchanged.
//
// Differences, all on purpose:
//   * buffers are page-locked (cudaHostAlloc) when a CUDA device is present, so the
//     synthesis result is copied device->host straight into data8/data16;
//   * fifo_enqueue links at the tail AND advances it (the reference forgets to move
//     fifo_tail, fifo.c:163-168, which silently drops buffers 1..6 of a run);
//     fifo_set_compat_drop(true) restores that loss for byte-identical iqdata.bin (guaranteed for the start-up
//     loss of buffers 1..6 with the reference's 8-buffer geometry; like the stock program, which later buffers are
//     lost once the consumer lags depends on timing);
//   * waits are `while` loops (the reference uses `if`, fifo.c:132-136,178-181).
#include <cuda_runtime_api.h>

#include <atomic>
#include <cassert>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpsb200.h"

namespace {

std::mutex g_mu;
std::condition_variable g_notempty, g_empty, g_free, g_full;
iq_buf *g_head = nullptr, *g_tail = nullptr, *g_freelist = nullptr;
bool g_halted = false;
bool g_compat_drop = false;
bool g_full_signalled = false;
std::vector<void *> g_pinned_ptrs;      // which sample buffers are page-locked (the rest came from calloc); under g_mu

void *alloc_bytes(size_t n) {
    void *p = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 &&
        cudaHostAlloc(&p, n, cudaHostAllocPortable) == cudaSuccess) {
        g_pinned_ptrs.push_back(p);
        memset(p, 0, n);
        return p;
    }
    cudaGetLastError();                  // no device, or the page-locked pool is exhausted: ordinary memory
    return calloc(1, n);
}

void free_bytes(void *p) {
    if (!p) return;
    for (size_t i = 0; i < g_pinned_ptrs.size(); i++)
        if (g_pinned_ptrs[i] == p) {
            g_pinned_ptrs[i] = g_pinned_ptrs.back();
            g_pinned_ptrs.pop_back();
            cudaFreeHost(p);
            return;
        }
    free(p);
}

void free_list(iq_buf *h) {
    while (h) {
        iq_buf *n = h->next;
        free_bytes(h->data8);
        free_bytes(h->data16);
        free(h);
        h = n;
    }
}

iq_buf *g_partial = nullptr;          // buffer being filled by gpsb200_fifo_push across calls

std::thread g_writer;
std::atomic<bool> g_writer_exit{false};
std::string g_path;
int g_sample_size = GPSB200_SC08;

}  // namespace

extern "C" {

void fifo_set_compat_drop(bool on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_compat_drop = on;
}

bool fifo_create(unsigned buffer_count, unsigned buffer_size, unsigned sample_size) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_halted = false;
    g_full_signalled = false;
    for (unsigned i = 0; i < buffer_count; i++) {
        iq_buf *b = (iq_buf *) calloc(1, sizeof(iq_buf));
        if (b) {
            if (sample_size == sizeof(signed short)) b->data16 = (signed short *) alloc_bytes((size_t) buffer_size * 2);
            else b->data8 = (signed char *) alloc_bytes((size_t) buffer_size);
        }
        if (!b || (!b->data8 && !b->data16)) {         // release what was created so far (fifo.c:60 calls fifo_destroy)
            free(b);
            free_list(g_freelist);
            g_freelist = nullptr;
            return false;
        }
        b->totalLength = buffer_size;
        b->next = g_freelist;
        g_freelist = b;
    }
    return true;
}

void fifo_destroy(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    free_list(g_head);
    free_list(g_freelist);
    if (g_partial) {
        g_partial->next = nullptr;
        free_list(g_partial);
    }
    g_head = g_tail = g_freelist = g_partial = nullptr;
}

void fifo_wait_next(void) {
    std::unique_lock<std::mutex> lk(g_mu);
    g_empty.wait(lk, [] { return !g_head || g_halted; });
}

void fifo_wait_full(void) {
    std::unique_lock<std::mutex> lk(g_mu);
    g_full.wait(lk, [] { return g_full_signalled || g_halted; });
}

void fifo_halt(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    while (g_head) {
        iq_buf *b = g_head;
        g_head = b->next;
        b->next = g_freelist;
        g_freelist = b;
    }
    g_tail = nullptr;
    g_halted = true;
    g_notempty.notify_all();
    g_empty.notify_all();
    g_free.notify_all();
    g_full.notify_all();
}

struct iq_buf *fifo_acquire(void) {
    std::unique_lock<std::mutex> lk(g_mu);
    if (!g_halted && !g_freelist) {
        g_full_signalled = true;          // every buffer is queued: "full" (fifo.c:132-133)
        g_full.notify_all();
    }
    g_free.wait(lk, [] { return g_freelist || g_halted; });
    if (g_halted) return nullptr;
    iq_buf *r = g_freelist;
    g_freelist = r->next;
    r->validLength = 0;
    r->next = nullptr;
    return r;
}

void fifo_enqueue(struct iq_buf *buf) {
    assert(buf->validLength <= buf->totalLength);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_halted) {
        buf->next = g_freelist;
        g_freelist = buf;
        return;
    }
    buf->next = nullptr;
    if (!g_head) {
        g_head = g_tail = buf;
        g_notempty.notify_one();
    } else if (g_compat_drop) {
        // stock behaviour: overwrite tail->next, never advance the tail; the buffer that
        // was linked there before is leaked (fifo.c:167)
        g_tail->next = buf;
    } else {
        g_tail->next = buf;
        g_tail = buf;
    }
}

struct iq_buf *fifo_dequeue(void) {
    std::unique_lock<std::mutex> lk(g_mu);
    g_notempty.wait(lk, [] { return g_head || g_halted; });
    if (g_halted) return nullptr;
    iq_buf *r = g_head;
    g_head = r->next;
    r->next = nullptr;
    if (!g_head) {
        g_tail = nullptr;
        g_empty.notify_all();
    }
    return r;
}

void fifo_release(struct iq_buf *buf) {
    std::lock_guard<std::mutex> lk(g_mu);
    buf->next = g_freelist;
    g_freelist = buf;
    g_free.notify_one();
}

// ---- stream -> FIFO buffers of any size (the HackRF cadence, gps.c:2847-2856) -------------
// The HackRF sink consumes 262144-element buffers that do not align with the 600000-element
// blocks: the reference keeps filling one acquired buffer across block boundaries and enqueues
// it whenever it is full. Same behaviour for a contiguous stream of elements; the partly
// filled buffer is kept for the next call (or flushed by fifo_push_flush at the end).
int gpsb200_fifo_push(const void *elems, size_t count, int sample_size) {
    const char *src = (const char *) elems;
    const size_t es = sample_size == GPSB200_SC16 ? 2 : 1;
    while (count > 0) {
        if (!g_partial) {
            g_partial = fifo_acquire();
            if (!g_partial) return GPSB200_ERR_ARG;          // halted
        }
        iq_buf *b = g_partial;
        const size_t room = b->totalLength - b->validLength;
        const size_t n = count < room ? count : room;
        char *dst = es == 2 ? (char *) b->data16 : (char *) b->data8;
        if (!dst) return GPSB200_ERR_ARG;                     // FIFO created for the other sample size
        memcpy(dst + (size_t) b->validLength * es, src, n * es);
        b->validLength += (unsigned) n;
        src += n * es;
        count -= n;
        if (b->validLength == b->totalLength) {
            fifo_enqueue(b);
            g_partial = nullptr;
        }
    }
    return GPSB200_OK;
}

int gpsb200_fifo_push_flush(void) {
    if (g_partial) {
        fifo_enqueue(g_partial);                              // validLength < totalLength: a short last buffer
        g_partial = nullptr;
    }
    return GPSB200_OK;
}

// ---- iqfile sink (sdr_iqfile.c:22-77): dequeue -> fwrite -> release -------------------
int gpsb200_iqfile_start(const char *path, int sample_size) {
    if (g_writer.joinable()) return GPSB200_ERR_ARG;
    g_path = path ? path : "iqdata.bin";            // sdr_iqfile.c:24
    g_sample_size = sample_size;
    FILE *fp = fopen(g_path.c_str(), "wb");
    if (!fp) return GPSB200_ERR_ARG;
    g_writer_exit = false;
    g_writer = std::thread([fp] {
        while (!g_writer_exit) {
            iq_buf *iq = fifo_dequeue();
            if (!iq) break;
            if (g_sample_size == GPSB200_SC16) fwrite(iq->data16, 2, iq->validLength, fp);
            else fwrite(iq->data8, 1, iq->validLength, fp);
            fifo_release(iq);
        }
        fclose(fp);
    });
    return GPSB200_OK;
}

void gpsb200_iqfile_stop(void) {
    if (!g_writer.joinable()) return;
    fifo_wait_next();              // drain what is queued (the reference drops it, sdr_iqfile.c:66-71)
    g_writer_exit = true;
    fifo_halt();
    g_writer.join();
}

}  // extern "C"
