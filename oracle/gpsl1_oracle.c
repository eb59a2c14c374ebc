/* TEST INFRASTRUCTURE -- parity oracle, not product code. See gpsl1_oracle.h.
 *
 * Written from the semantics in SURVEY.md section 8a, channel-major (the
 * reference is sample-major; integer accumulation commutes), with the
 * double-precision recurrences evaluated exactly as the reference does:
 * separate multiply and add, round-to-nearest, one step per sample
 * (compile with -ffp-contract=off). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gpsl1_oracle.h"

/* First quadrant of the reference sine table: sinTable512[k], k = 0..127
 * (gps.c:145-178). Equal to floor(250*sin(2*pi*(k+0.5)/512)+0.5) except k = 35,
 * where the reference holds 105 (the formula gives 105.50007 -> 106). The other
 * three quadrants follow from sin[255-k] = sin[k], sin[k+256] = -sin[k], and
 * cosTable512[k] = sinTable512[(k+128)&511] (gps.c:180-213); checked against
 * the reference arrays by tests/test_oracle.py via the golden TABLES record. */
static const uint8_t quarter_sine[128] = {
      2,   5,   8,  11,  14,  17,  20,  23,  26,  29,  32,  35,  38,  41,  44,  47,
     50,  53,  56,  59,  62,  65,  68,  71,  74,  77,  80,  83,  86,  89,  91,  94,
     97, 100, 103, 105, 108, 111, 114, 116, 119, 122, 125, 127, 130, 132, 135, 138,
    140, 143, 145, 148, 150, 153, 155, 157, 160, 162, 164, 167, 169, 171, 173, 176,
    178, 180, 182, 184, 186, 188, 190, 192, 194, 196, 198, 200, 202, 204, 205, 207,
    209, 210, 212, 214, 215, 217, 218, 220, 221, 223, 224, 225, 227, 228, 229, 230,
    232, 233, 234, 235, 236, 237, 238, 239, 240, 241, 241, 242, 243, 244, 244, 245,
    245, 246, 247, 247, 248, 248, 248, 249, 249, 249, 249, 250, 250, 250, 250, 250
};

static int sine_at(int k) {
    k &= 511;
    int q = k & 255;
    int v = quarter_sine[q < 128 ? q : 255 - q];
    return k < 256 ? v : -v;
}

void oracle_tables(int32_t sin512[512], int32_t cos512[512]) {
    for (int k = 0; k < 512; k++) {
        sin512[k] = sine_at(k);
        cos512[k] = sine_at(k + 128);
    }
}

/* G2 output delay (chips) per PRN, IS-GPS-200 Table 3-Ia (gps.c:273-278). */
static const uint16_t g2_delay[32] = {
      5,   6,   7,   8,  17,  18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,
    469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862
};

int oracle_codegen(int prn, uint8_t ca[ORACLE_CA_LEN]) {
    if (prn < 1 || prn > 32) return -1;
    uint8_t g1[ORACLE_CA_LEN], g2[ORACLE_CA_LEN];
    unsigned r1 = 0x3FF, r2 = 0x3FF;           /* stage s is bit s-1; all ones */
    for (int i = 0; i < ORACLE_CA_LEN; i++) {
        g1[i] = (r1 >> 9) & 1;
        g2[i] = (r2 >> 9) & 1;
        unsigned f1 = ((r1 >> 2) ^ (r1 >> 9)) & 1;                         /* x^10 + x^3 + 1 */
        unsigned f2 = ((r2 >> 1) ^ (r2 >> 2) ^ (r2 >> 5) ^ (r2 >> 7) ^ (r2 >> 8) ^ (r2 >> 9)) & 1;
        r1 = ((r1 << 1) | f1) & 0x3FF;
        r2 = ((r2 << 1) | f2) & 0x3FF;
    }
    int d = g2_delay[prn - 1];
    for (int i = 0; i < ORACLE_CA_LEN; i++)
        ca[i] = g1[i] ^ g2[(i + ORACLE_CA_LEN - d) % ORACLE_CA_LEN];
    return 0;
}

void oracle_synth_block(oracle_chan_t *ch, int nchan, int nsamp, int16_t *iq16) {
    const double delt = 1.0 / 3000000.0;       /* gps.c:2298, sdr.h:21 */
    int32_t *acc = calloc((size_t) 2 * nsamp, sizeof *acc);
    int32_t s512[512], c512[512];
    oracle_tables(s512, c512);

    for (int c = 0; c < nchan; c++) {
        oracle_chan_t *p = &ch[c];
        if (p->prn <= 0) continue;
        uint8_t ca[ORACLE_CA_LEN];
        oracle_codegen(p->prn, ca);
        /* gps.c:2781-2782: int product -> double, times gain, truncated to int.
         * The +-1 factors commute with the truncation, so tabulate |.| per k. */
        int32_t ai[512], aq[512];
        for (int k = 0; k < 512; k++) {
            ai[k] = (int32_t) ((double) c512[k] * p->gain);
            aq[k] = (int32_t) ((double) s512[k] * p->gain);
        }
        const double dcarr = p->f_carr * delt; /* gps.c:2821, rounded once */
        const double dcode = p->f_code * delt; /* gps.c:2789 */
        double carr = p->carr_phase, code = p->code_phase;
        int iword = p->iword, ibit = p->ibit, icode = p->icode;
        int bit = (int) ((p->dwrd[iword] >> (29 - ibit)) & 1u);  /* gps.c:2060 */
        for (int n = 0; n < nsamp; n++) {
            int k = (int) floor(carr * 512.0);                   /* gps.c:2775 */
            /* dataBit*codeCA, both mapped 0/1 -> -1/+1: product is +1 iff bit == chip */
            int sign = (bit == ca[(int) code]) ? 1 : -1;
            acc[2 * n] += sign * ai[k];
            acc[2 * n + 1] += sign * aq[k];

            code += dcode;                                       /* gps.c:2789-2817 */
            if (code >= 1023.0) {
                code -= 1023.0;
                if (++icode >= 20) {
                    icode = 0;
                    if (++ibit >= 30) { ibit = 0; iword++; }
                    bit = (int) ((p->dwrd[iword] >> (29 - ibit)) & 1u);
                }
            }
            carr += dcarr;                                       /* gps.c:2821-2826 */
            if (carr >= 1.0) carr -= 1.0;
            else if (carr < 0.0) {
                carr += 1.0;
                /* The reference can reach carr == 1.0 here (tiny negative + 1.0 rounds up) and
                 * then indexes its 512-entry tables with 512: undefined behaviour (SURVEY.md
                 * hard part 6). Oracle and product both clamp to the largest double below 1. */
                if (carr >= 1.0) carr = 0.99999999999999988897769753748434595763683319091796875;
            }
        }
        p->carr_phase = carr; p->code_phase = code;
        p->iword = iword; p->ibit = ibit; p->icode = icode;
    }
    for (int j = 0; j < 2 * nsamp; j++) iq16[j] = (int16_t) acc[j]; /* gps.c:2834-2835 */
    free(acc);
}

void oracle_quantize8(const int16_t *iq16, int nelem, int8_t *iq8) {
    for (int j = 0; j < nelem; j++)
        iq8[j] = (int8_t) (uint8_t) ((iq16[j] >> 4) & 0xFF);     /* gps.c:2844 */
}
