/* TEST INFRASTRUCTURE -- not product code.
 *
 * Drop-in check of the FIFO boundary against the reference's OWN consumer code: the reference's unmodified sdr.c
 * and sdr_iqfile.c (dispatch table sdr.c:35-87, writer thread sdr_iqfile.c:22-77) are compiled as they are and
 * linked with libgpsb200.so INSTEAD OF fifo.o. This main plays the producer exactly like gps_thread_ep does
 * (fifo_acquire -> fill 600000 elements -> fifo_enqueue, gps.c:2698,2839-2865): it replays a recorded IQ stream.
 * The sink writes ./iqdata.bin (sdr_iqfile.c:24), which must equal the input. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <pthread.h>
#include "gps-sim.h"
#include "sdr.h"
#include "fifo.h"
#include "gui.h"

static simulator_t simulator;

void set_thread_name(const char *name) { (void) name; }
int thread_to_core(int core_id) { (void) core_id; return 0; }

static FILE *g_in;
static int g_blocks;

/* the producer: what gps_thread_ep does with the FIFO (gps.c:2698, 2839-2865), replaying a recorded stream */
static void *producer(void *arg) {
    (void) arg;
    const size_t esz = simulator.sample_size == SC16 ? 2 : 1;
    struct iq_buf *iq = fifo_acquire();
    while (iq) {
        void *dst = simulator.sample_size == SC16 ? (void *) iq->data16 : (void *) iq->data8;
        if (fread(dst, esz, IQ_BUFFER_SIZE, g_in) != IQ_BUFFER_SIZE) break;      /* whole blocks only */
        iq->validLength = IQ_BUFFER_SIZE;
        fifo_enqueue(iq);
        g_blocks++;
        simulator.gps_thread_running = true;
        iq = fifo_acquire();
    }
    simulator.gps_thread_exit = true;
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "ref_sinkfeed STREAM.bin 8|16\n"); return 2; }
    simulator.sample_size = atoi(argv[2]) == 16 ? SC16 : SC08;
    simulator.sdr_name = "iqfile";
    g_in = fopen(argv[1], "rb");
    if (!g_in) { perror(argv[1]); return 1; }
    if (sdr_init(&simulator) != 0) return 1;                       /* -> sdr_iqfile_init -> fifo_create (ours) */
    pthread_t th;
    pthread_create(&th, NULL, producer, NULL);
    while (!simulator.gps_thread_running && !simulator.gps_thread_exit) usleep(1000);
    sdr_run();                                                     /* gps-sim.c:325: waits for the primed FIFO, starts the writer */
    pthread_join(th, NULL);
    fifo_wait_next();                                              /* let the writer drain what is queued ... */
    usleep(300000);                                                /* ... and finish its last fwrite before fifo_halt() */
    sdr_close();
    fclose(g_in);
    printf("{\"blocks\": %d}\n", g_blocks);
    return 0;
}
