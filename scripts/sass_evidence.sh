#!/bin/bash
# SASS evidence for the tcgen05 / TMA claims: per kernel of the in-tree objects, the count of the Blackwell mnemonics
# (UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,
#  FFMA2 / FMUL2 = packed f32x2, MUFU = special-function unit).  Usage: bash scripts/sass_evidence.sh > profiles/r02_sass.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass of uncrtaints_b200/_build/*.o (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a)"
for o in gemm_tc dwconv_rows norm temporal ltae_v head_loss inconv se optim metrics gemm_simt; do
  f=uncrtaints_b200/_build/$o.o
  [ -f $f ] || continue
  cuobjdump -sass $f | awk -v obj=$o '
    /Function :/ { if (name != "") print_row(); cmd="echo " $3 " | c++filt"; cmd | getline name; close(cmd); gsub(/ub::\(anonymous namespace\)::/,"",name); gsub(/ub::tc::/,"",name); gsub(/ub::/,"",name); sub(/\(.*/,"",name); delete c; n=0 }
    { n++ }
    /UTCHMMA/ {c["UTCHMMA"]++} /LDTM/ {c["LDTM"]++} /UTCBAR/ {c["UTCBAR"]++} /UBLKCP/ {c["UBLKCP"]++} /UTMALDG|UTMASTG/ {c["UTMA"]++}
    /SYNCS/ {c["SYNCS"]++} /FFMA2|FMUL2/ {c["F32x2"]++} /MUFU/ {c["MUFU"]++} /ATOMG|RED\./ {c["ATOM"]++} /HMMA\./ {c["HMMA"]++}
    function print_row() { printf "%-14s %-110s lines=%-6d UTCHMMA=%-4d LDTM=%-3d UTCBAR=%-3d UBLKCP=%-3d UTMA=%-2d SYNCS=%-3d F32x2=%-4d MUFU=%-3d\n", obj, substr(name,1,110), n, c["UTCHMMA"], c["LDTM"], c["UTCBAR"], c["UBLKCP"], c["UTMA"], c["SYNCS"], c["F32x2"], c["MUFU"] }
    END { if (name != "") print_row() }'
done
