// gpsb200-sim: file-sink driver with the reference's command-line vocabulary (help.h:20-53:
// -e nav file, -l location, -d duration, -m motion file, -s start, --iq16, -I no ionosphere).
// RINEX + location -> scenario engine (host) -> CUDA synthesis -> reference-compatible FIFO ->
// iqfile writer. Output is byte-identical to the reference's enqueue stream; --compat-drop
// reproduces the stock program's iqdata.bin (which lacks blocks 1..6, fifo.c:163-168).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "../include/gpsb200.h"

static void usage() {
    fprintf(stderr,
            "gpsb200-sim -e NAV [-3] -l lat,lon,h [-d SEC] [-m motion.csv] [-s y/m/d,h:m:s] [--iq16] [-I]\n"
            "            [--chan N] [-o iqdata.bin] [--compat-drop]\n");
    exit(2);
}

int main(int argc, char **argv) {
    gpsb200_scenario_config_t sc{};
    sc.ionosphere_enable = 1;
    sc.max_chan = 12;
    double dur = 300.0;
    int sample_size = GPSB200_SC08;
    bool compat = false;
    std::string out = "iqdata.bin";
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&]() -> const char * {
            if (i + 1 >= argc) usage();
            return argv[++i];
        };
        if (a == "-e") sc.nav_file = need();
        else if (a == "-l") sscanf(need(), "%lf,%lf,%lf", &sc.lat_deg, &sc.lon_deg, &sc.height_m);
        else if (a == "-d") dur = atof(need());
        else if (a == "-m") sc.motion_file = need();
        else if (a == "-s")
            sscanf(need(), "%d/%d/%d,%d:%d:%lf", &sc.start_year, &sc.start_month, &sc.start_day, &sc.start_hour,
                   &sc.start_min, &sc.start_sec);
        else if (a == "--iq16") sample_size = GPSB200_SC16;
        else if (a == "-I") sc.ionosphere_enable = 0;
        else if (a == "-3") sc.rinex3 = 1;
        else if (a == "--chan") sc.max_chan = atoi(need());
        else if (a == "-o") out = need();
        else if (a == "--compat-drop") compat = true;
        else usage();
    }
    if (!sc.nav_file) usage();
    sc.duration_ds = (int) (dur * 10.0 + 0.5);                  // gps-sim.c:140

    gpsb200_scenario_t *scn = nullptr;
    if (gpsb200_scenario_create(&sc, &scn) != GPSB200_OK) {
        fprintf(stderr, "scenario: %s\n", gpsb200_scenario_error(scn));
        return 1;
    }
    const int nblk = gpsb200_scenario_blocks(scn), nchan = gpsb200_scenario_channels(scn);
    const int nframes = gpsb200_scenario_nav_frames(scn);
    const gpsb200_chan_t *chans = gpsb200_scenario_chans(scn);
    const uint32_t *nav = gpsb200_scenario_nav(scn);

    const int batch = 256;
    gpsb200_config_t cfg{};
    cfg.max_chan = nchan;
    cfg.max_blocks = batch;
    cfg.max_nav_frames = nframes;
    gpsb200_ctx_t *ctx = nullptr;
    gpsb200_bind_numa(cfg.device);      // threads and pinned FIFO buffers next to the GPU (cf. thread_to_core, gps.c:2377)
    if (gpsb200_create(&cfg, &ctx) != GPSB200_OK) {
        fprintf(stderr, "gpsb200: %s\n", gpsb200_last_error(ctx));
        return 1;
    }
    for (int f = 0; f < nframes; f++)
        for (int c = 0; c < nchan; c++) gpsb200_set_nav(ctx, f, c, nav + ((size_t) f * nchan + c) * GPSB200_NAV_WORDS);

    fifo_set_compat_drop(compat);
    if (!fifo_create(8, GPSB200_BLOCK_ELEMS, sample_size)) return 1;       // sdr_iqfile.c:59, sdr.h:24
    void *stage = nullptr;
    const size_t blk_bytes = (size_t) GPSB200_BLOCK_ELEMS * sample_size;
    if (cudaHostAlloc(&stage, blk_bytes * batch, cudaHostAllocDefault) != cudaSuccess) return 1;

    std::vector<double> carr(nchan, 0.0);
    std::vector<gpsb200_chan_t> part;
    bool writer = false;
    int queued = 0;
    for (int b0 = 0; b0 < nblk; b0 += batch) {
        const int nb = std::min(batch, nblk - b0);
        part.assign(chans + (size_t) b0 * nchan, chans + (size_t) (b0 + nb) * nchan);
        if (b0 > 0)                                             // continue the carrier chain across calls
            for (int c = 0; c < nchan; c++)
                if (part[c].prn > 0 && part[c].prn == chans[(size_t) (b0 - 1) * nchan + c].prn) part[c].carr_phase = carr[c];
        if (gpsb200_synth_blocks(ctx, part.data(), nb, nchan, sample_size, stage, carr.data(), nullptr) != GPSB200_OK) {
            fprintf(stderr, "gpsb200: %s\n", gpsb200_last_error(ctx));
            return 1;
        }
        for (int b = 0; b < nb; b++) {
            // like the reference (sdr_iqfile.c:73-77) the writer starts once the FIFO is full (or the run ends)
            if (!writer && queued == 8) {
                if (gpsb200_iqfile_start(out.c_str(), sample_size) != GPSB200_OK) return 1;
                writer = true;
            }
            struct iq_buf *iq = fifo_acquire();
            if (!iq) return 1;
            memcpy(sample_size == GPSB200_SC16 ? (void *) iq->data16 : (void *) iq->data8,
                   (char *) stage + (size_t) b * blk_bytes, blk_bytes);
            iq->validLength = GPSB200_BLOCK_ELEMS;
            fifo_enqueue(iq);
            queued++;
        }
    }
    if (!writer && gpsb200_iqfile_start(out.c_str(), sample_size) != GPSB200_OK) return 1;
    gpsb200_iqfile_stop();
    fifo_destroy();
    cudaFreeHost(stage);
    gpsb200_destroy(ctx);
    gpsb200_scenario_destroy(scn);
    fprintf(stderr, "gpsb200-sim: %d blocks (%d channels) -> %s\n", nblk, nchan, out.c_str());
    return 0;
}
