#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: apply the drop-in integration of INTEGRATION.md section 1 to a TEMPORARY copy of the
reference's gps.c (nothing of the reference is stored in this repository: the edit is described by line ranges of
reference @56b8776 and the inserted code lives in integration_*.inc next to this script).

    apply_integration.py /root/reference/gps.c OUT.c

Edits (1-based line numbers of the unmodified file):
    after 2320            #include "integration_decl.inc"       context pointer + carried carrier phases
    after 2698            #include "integration_setup.inc"      gpsb200_create
    replace 2767..2857    #include "integration_block.inc"      the sample loop and the quantise/pack loop
    after 2940            #include "integration_teardown.inc"   gpsb200_destroy
The HackRF cadence (262144-element buffers, gps.c:2847-2856) is not part of this minimal patch: use
gpsb200_fifo_push() for it."""
import hashlib
import sys

SHA256_16 = "0a9f5a9a2fd28bd2"     # gps.c of Mictronics/multi-sdr-gps-sim @56b8776
ANCHORS = {2320: "short *iq_buff = NULL;", 2698: "struct iq_buf *iq = fifo_acquire();",
           2767: "for (isamp = 0; isamp < NUM_IQ_SAMPLES; isamp++) {", 2857: "}", 2940: "end_gps_thread:"}


def main(src, dst):
    raw = open(src, "rb").read()
    if hashlib.sha256(raw).hexdigest()[:16] != SHA256_16:
        sys.exit("apply_integration: %s is not gps.c @56b8776; line ranges would not fit" % src)
    lines = raw.decode("utf-8", "replace").split("\n")
    for no, text in ANCHORS.items():
        if lines[no - 1].strip() != text:
            sys.exit("apply_integration: line %d is %r, expected %r" % (no, lines[no - 1].strip(), text))
    out = []
    for no, line in enumerate(lines, 1):
        if 2767 <= no <= 2857:
            if no == 2767:
                out.append('#include "integration_block.inc"')
            continue
        out.append(line)
        if no == 2320:
            out.append('#include "integration_decl.inc"')
        elif no == 2698:
            out.append('#include "integration_setup.inc"')
        elif no == 2940:
            out.append('#include "integration_teardown.inc"')
    head = '#include "gpsb200.h"   /* drop-in integration, see oracle/ref_harness/apply_integration.py */\n'
    open(dst, "w").write(head + "\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
