#!/bin/bash
# static SASS instruction mix of the general and the FAST instances of the module stencil bodies (no GPU needed): nvcc cross-compile + cuobjdump
set -eu
cd "$(dirname "$0")/.."
tmp=$(mktemp -d)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -diag-suppress 550 -cubin -o "$tmp/f.cubin" scripts/sassmix/fast_instances.cu
echo "# static SASS of one cell's stencil body, general vs FAST instance (sm_100a, -fmad=false; out-of-line helpers -- tc_coefficient, tf_T4, tf_D, pv_point -- are counted once"
echo "# although a cell executes them 5 / 4 / 16 / 5 times, so the dynamic ratio is larger than the static one for tf and pv).  scripts/fast_instances_sass.sh"
cuobjdump -sass "$tmp/f.cubin" | awk '/Function : /{name=$3} / \/\*[0-9a-f]+\*\/ +[A-Z@]/{cnt[name]++; if ($0 ~ / D(ADD|MUL|FMA|SETP|MNMX)/) f[name]++; if ($0 ~ /MUFU/) m[name]++; if ($0 ~ / (IMAD|IADD3|IADD|ISETP|LEA|SEL|IMNMX|VIMNMX|VIADD|LOP3|SHF)/) ii[name]++; if ($0 ~ / LDG/) l[name]++; if ($0 ~ / LDC/) c[name]++; if ($0 ~ / (BRA|CALL|RET|BSSY|BSYNC)/) b[name]++} END{for (n in cnt) if (n ~ /^_Z[0-9]+(tc|tf|pv)_(fast|general)/) { s = n; sub(/^_Z[0-9]+/, "", s); sub(/N6spruce.*/, "", s); printf "%-12s total %5d  fp64 %4d  integer %4d  branch %4d  ldg %3d  ldc %3d  mufu %2d\n", s, cnt[n], f[n], ii[n], b[n], l[n], c[n], m[n]} }' | sort
rm -rf "$tmp"
