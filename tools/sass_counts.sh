#!/bin/bash
# Blackwell-native evidence (run HERE, no GPU): per kernel of libxvec_b200.so the count of tcgen05 / TMEM / TMA SASS
# instructions.  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit.
#   bash tools/sass_counts.sh > profiles/r02_sass_counts.txt
LIB=${1:-x-vector-kaldi-tf_b200/libxvec_b200.so}
echo "# cuobjdump -sass $LIB | per-function counts (sm_100a); $(date -u +%Y-%m-%dT%H:%MZ); $(nvcc --version | tail -1)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3 }
  /UTCHMMA/ { a[fn]++ } /LDTM/ { b[fn]++ } /UTMALDG/ { c[fn]++ } /UTMASTG/ { d[fn]++ } /UTCBAR/ { e[fn]++ } /UTMAPF|UTMACCTL/ { f[fn]++ }
  /Function :/ { seen[fn]=1 }
  END { printf "%-8s %-6s %-8s %-8s %-7s %s\n", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "function";
        for (k in seen) if (a[k]+b[k]+c[k]+d[k]+e[k] > 0) printf "%-8d %-6d %-8d %-8d %-7d %s\n", a[k], b[k], c[k], d[k], e[k], k }' | (read -r hdr; echo "$hdr"; sort -k6 | c++filt)
echo "# totals:"
cuobjdump -sass "$LIB" | grep -oE "UTCHMMA[.A-Z0-9_]*|LDTM[.A-Z0-9x_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*" | sort | uniq -c | sort -rn
