#!/usr/bin/env bash
# SASS listings of the two tensor-core kernels as built in-tree (no GPU needed):
#   tools/sass_listing.sh      -> profiles/sass_tc_pass1.txt, profiles/sass_tc_exact.txt
# Each file: the nvcc command line, a mnemonic histogram (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
# UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg), then the listing
# with the instruction encodings stripped.
set -euo pipefail
cd "$(dirname "$0")/.."
python -m optimalmodulationds_b200.build > /dev/null
list() {   # object, start pattern, stop pattern, output
  {
    echo "# $(git rev-parse --short HEAD 2>/dev/null || echo '?') :: cuobjdump -sass optimalmodulationds_b200/build/$1, function matching /$2/"
    echo "# nvcc $(python -c 'from optimalmodulationds_b200 import build as b; print(" ".join(b.NVCC_FLAGS[:8]))')"
    cuobjdump -sass "optimalmodulationds_b200/build/$1" | awk "/Function : .*$2/{f=1} f&&/Function : /&&!/$2/{f=0} f" \
      | grep -v '^\s*$' | sed -E 's@/\* 0x[0-9a-f]+ \*/@@; s/[[:space:]]+$//' > /tmp/sass_body.txt
    echo "# mnemonic histogram"
    grep -oE '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z][A-Z0-9_.]+' /tmp/sass_body.txt | awk '{print $NF}' | sed -E 's/\..*//' \
      | sort | uniq -c | sort -rn | awk '{printf "#   %-14s %s\n", $2, $1}'
    cat /tmp/sass_body.txt
  } > "$4"
  echo "$4: $(wc -l < "$4") lines"
}
list tc_pass1.o 'tc_pass1_kernelILb0ELb1ELb1' x profiles/sass_tc_pass1.txt
list tc_exact.o 'tc_exact_kernelILi1ELb0' x profiles/sass_tc_exact.txt
