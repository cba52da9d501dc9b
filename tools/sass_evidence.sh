#!/usr/bin/env bash
# SASS-level proof that the product kernels are tcgen05/TMA code (no GPU needed):
# per kernel, counts of the Blackwell mnemonics named in /opt/skills/guides/B200_PROFILING.md.
set -eu
LIB=${1:-nnaudio_b200/libnnab.so}
printf "%-96s %8s %8s %6s %7s %6s %8s %6s\n" "kernel (mangled; <BK, STAGES, FMT>)" UTCHMMA UTMALDG LDTM UTCBAR SYNCS RED/ATOM FFMA
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; next }
  { for (m in pat) if ($0 ~ pat[m]) cnt[fn, m]++ }
  BEGIN { pat["UTCHMMA"]="UTCHMMA"; pat["UTMALDG"]="UTMALDG"; pat["LDTM"]="LDTM"; pat["UTCBAR"]="UTCBAR";
          pat["SYNCS"]="SYNCS"; pat["UTCATOMSWS"]="UTCATOMSWS"; pat["HMMA"]="[^C]HMMA"; pat["FFMA"]="FFMA";
          pat["RED/ATOM"]="(RED|ATOMG)" }
  END {
    for (k in cnt) { split(k, a, SUBSEP); fns[a[1]]=1 }
    for (f in fns) if ((f, "UTCHMMA") in cnt || (f, "UTMALDG") in cnt)
      printf "%-96s %8d %8d %6d %7d %6d %8d %6d\n", substr(f, 1, 96), cnt[f,"UTCHMMA"], cnt[f,"UTMALDG"], cnt[f,"LDTM"], cnt[f,"UTCBAR"], cnt[f,"SYNCS"], cnt[f,"RED/ATOM"], cnt[f,"FFMA"];
  }' | sort
echo
echo "# registers / stack per tensor-core kernel (cuobjdump -res-usage; dynamic smem is set at launch: 197.9 KB)"
cuobjdump -res-usage "$LIB" 2>/dev/null | grep -A1 -E "Function _ZN4nnab(1[6-9]framed_tc|13fir_tc_kernel)" | grep -v "^--" | paste - - \
  | sed -E 's/.*Function (_ZN4nnab[0-9]+[a-z_0-9]+(ILi[0-9]+E?(Li[0-9]+E?)*E)?)[^:]*:\s*/\1  /; s/ SHARED.*//' | sort
