// TSan harness for csrc/fifo.cpp: producer/consumer threads through the C API
#include <cstdio>
#include <thread>
#include <vector>
#include "gpsb200.h"
int main() {
    for (int round = 0; round < 3; round++) {
        if (!fifo_create(8, 4096, round == 1 ? 2 : 1)) return 1;
        long sum_in = 0, sum_out = 0; int nout = 0;
        std::thread cons([&] { while (true) { iq_buf *b = fifo_dequeue(); if (!b) break; sum_out += b->data8 ? b->data8[0] : b->data16[0]; nout++; fifo_release(b);} });
        std::thread prod([&] { for (int i = 0; i < 2000; i++) { iq_buf *b = fifo_acquire(); if (!b) break; if (b->data8) b->data8[0] = (signed char)(i & 63); else b->data16[0] = (short)(i & 63); b->validLength = 4096; sum_in += i & 63; fifo_enqueue(b);} });
        prod.join(); fifo_wait_next(); fifo_halt(); cons.join(); fifo_destroy();
        printf("round %d: in %ld out %ld n %d\n", round, sum_in, sum_out, nout);
        if (sum_in != sum_out || nout != 2000) return 2;
    }
    return 0;
}
