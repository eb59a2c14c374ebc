/* TEST INFRASTRUCTURE -- not product code.
 *
 * Headless main() for the reference program AS SHIPPED: reference gps.c,
 * fifo.c, sdr.c, sdr_iqfile.c and almanac.c are compiled unmodified and this
 * file only replaces gps-sim.c (argp/ncurses UI, gps-sim.c:267-418). It follows
 * gps-sim.c's start-up order: sdr_init -> producer thread -> wait for
 * gps_init_done -> sdr_run -> (join producer) -> drain -> sdr_close.
 * Output: ./iqdata.bin (name fixed by sdr_iqfile.c:24). Because of the tail
 * bug at fifo.c:163-168 this file lacks blocks 1..6 of the enqueue stream;
 * tests/test_oracle_ref.py checks exactly that relation. */
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

int main(int argc, char **argv) {
    double dur = 10.0;
    simulator.ionosphere_enable = true;
    simulator.almanac_enable = false;
    simulator.sample_size = SC08;
    simulator.sdr_name = "iqfile";
    pthread_cond_init(&simulator.gps_init_done, NULL);
    pthread_mutex_init(&simulator.gps_lock, NULL);
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-e") && i + 1 < argc) simulator.nav_file_name = argv[++i];
        else if (!strcmp(argv[i], "-l") && i + 1 < argc)
            sscanf(argv[++i], "%lf,%lf,%lf", &simulator.location.lat, &simulator.location.lon,
                   &simulator.location.height);
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) dur = atof(argv[++i]);
        else if (!strcmp(argv[i], "-m") && i + 1 < argc) simulator.motion_file_name = argv[++i];
        else if (!strcmp(argv[i], "--iq16")) simulator.sample_size = SC16;
        else { fprintf(stderr, "ref_stock -e NAV -l lat,lon,h -d SEC [--iq16] [-m csv]\n"); return 2; }
    }
    if (!simulator.nav_file_name || dur < 0.9) return 2; /* < 9 blocks never fills the FIFO */
    simulator.duration = (int) (dur * 10.0 + 0.5);

    if (sdr_init(&simulator) != 0) return 1;
    pthread_create(&simulator.gps_thread, NULL, gps_thread_ep, &simulator);
    /* gps.c:2711 signals gps_init_done without holding gps_lock, so poll the
     * flags instead of risking a lost wake-up (gps-sim.c:316 uses a 30 s timed wait). */
    while (!simulator.gps_thread_running && !simulator.gps_thread_exit)
        usleep(1000);
    if (simulator.gps_thread_exit && !simulator.gps_thread_running) return 1;
    sdr_run();
    pthread_join(simulator.gps_thread, NULL);
    fifo_wait_next();  /* let the writer drain what is queued */
    usleep(300000);    /* ...and finish its last fwrite before fifo_halt() */
    sdr_close();
    return 0;
}
