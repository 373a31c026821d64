#!/bin/bash
# Per-kernel SASS opcode counts of the shipped library (run here, no GPU needed):
#   tools/sass_opcodes.sh > profiles/r02_sass_opcodes.txt
# The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md names as proof of Blackwell-native code:
#   UTCIMMA / UTCHMMA  tcgen05.mma     LDTM      tcgen05.ld (TMEM -> registers)     UTCBAR   tcgen05.commit
#   UTMALDG / UTMASTG  TMA tensor copy UBLKCP    cp.async.bulk (1-D bulk copy)      SYNCS    mbarrier
#   HMMA              mma.sync (legacy register-fragment MMA: the RGB relight kernel)
#   FFMA2 / FMUL2      packed f32x2    I2IP      cvt.pack.sat (saturating byte pack) VIMNMX   packed integer min/max
LIB=${1:-rti_b200/lib/libptm_b200.so}
echo "# $LIB  ($(git log -1 --format=%h 2>/dev/null))"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
/^[[:space:]]+\/\*[0-9a-f]{4}\*\// {
    op=$2; sub(/;$/,"",op); split(op, parts, "."); base=parts[1];
    if (base=="UTCIMMA"||base=="UTCHMMA"||base=="UTCQMMA"||base=="LDTM"||base=="STTM"||base=="UTCBAR"||base=="UTMALDG"||base=="UTMASTG"||base=="UBLKCP"||base=="SYNCS"||base=="FFMA2"||base=="FMUL2"||base=="FADD2"||base=="I2IP"||base=="VIMNMX"||base=="FFMA"||base=="ACQBULK"||base=="UTCATOMSWS"||base=="REDUX")
        n[fn" "base]++
    total[fn]++
}
END { for (k in n) { split(k, a, " "); print a[1], a[2], n[k], "of", total[a[1]] } }' | sort | c++filt | awk '{fn=$0; sub(/ [A-Z0-9]+ [0-9]+ of [0-9]+$/,"",fn); print}' >/dev/null
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
/^[[:space:]]+\/\*[0-9a-f]{4}\*\// {
    op=$2; sub(/;$/,"",op); split(op, parts, "."); base=parts[1];
    if (base ~ /^(HMMA|UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTCBAR|UTMALDG|UTMASTG|UBLKCP|SYNCS|FFMA2|FMUL2|FADD2|I2IP|VIMNMX|FFMA|ACQBULK|REDUX)$/)
        n[fn, base]++
    total[fn]++
}
END {
    for (fn in total) {
        line=""
        for (k in n) { split(k, a, SUBSEP); if (a[1]==fn) line=line" "a[2]"="n[k] }
        if (line != "") print fn" ["total[fn]" instructions]"line
    }
}' | sort | while read -r sym rest; do echo "$(echo "$sym" | c++filt | cut -c1-110) $rest"; done
