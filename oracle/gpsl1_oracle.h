/* TEST INFRASTRUCTURE -- parity oracle, not product code.
 *
 * Plain-C CPU restatement of the reference's GPS L1 C/A sample-synthesis path
 * (Mictronics/multi-sdr-gps-sim, gps.c:2767-2857 plus the tables gps.c:145-213
 * and codegen gps.c:272-309). Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (libgpsb200.so) never links or calls it.
 *
 * Pinning: the reference holds no tests or golden vectors,
 * so this restatement is pinned against the reference ITSELF: tests/golden/ holds
 * parameter dumps and output digests produced by oracle/_ref/ref_dump{12,32}
 * (the unmodified reference gps.c behind a recording FIFO), see
 * tests/golden/make_golden.py and tests/test_oracle.py.
 */
#ifndef GPSL1_ORACLE_H
#define GPSL1_ORACLE_H
#include <stdint.h>

#define ORACLE_SAMPLES_PER_BLOCK 300000  /* sdr.h:26 NUM_IQ_SAMPLES */
#define ORACLE_CA_LEN 1023               /* gps.h:57 */
#define ORACLE_NAV_WORDS 60              /* gps.h:52 N_DWRD */

/* Per-channel record: the fields of channel_t (gps.h:213-236) and gain[]
 * (gps.c:2300) that the sample loop reads. State fields are updated in place
 * exactly as the reference loop leaves them. */
typedef struct {
    int32_t prn;            /* 1..32, 0 = channel unused (gps.c:2772) */
    int32_t iword, ibit, icode;
    double f_carr, f_code;  /* Hz */
    double carr_phase;      /* cycles, [0,1)  -- persists across blocks */
    double code_phase;      /* chips, [0,1023) */
    double gain;
    uint32_t dwrd[ORACLE_NAV_WORDS]; /* 30-bit NAV words, MSB first */
} oracle_chan_t;

/* C/A Gold code of PRN 1..32 as 0/1 chips (gps.c:272-309). Returns 0, or -1. */
int oracle_codegen(int prn, uint8_t ca[ORACLE_CA_LEN]);
/* The two 512-entry carrier tables (gps.c:145-213). */
void oracle_tables(int32_t sin512[512], int32_t cos512[512]);
/* One block: nsamp complex samples, iq16[2*n] = I, iq16[2*n+1] = Q
 * (gps.c:2767-2836). Channel state advances. */
void oracle_synth_block(oracle_chan_t *ch, int nchan, int nsamp, int16_t *iq16);
/* gps.c:2839-2845: int16 -> int8 by arithmetic >>4 and modulo-256 narrowing. */
void oracle_quantize8(const int16_t *iq16, int nelem, int8_t *iq8);
#endif
