#!/bin/bash
# SASS evidence for profiles/: counts of the tensor-core / TMA / mbarrier / packed-math instructions per object file
# of spatial-vae_b200/csrc (built by __graft_entry__.build()), and the resources of the five big kernels.
cd "$(dirname "$0")/../spatial-vae_b200/csrc" || exit 1
echo "# SASS evidence, libsvae_b200.so built for sm_100a (cuobjdump -sass), instruction counts per source file"
for o in tc_gemm.o tc_bwd.o step_kernels.o; do
  echo "== $o"
  cuobjdump -sass $o | grep -oE "\b(UTC[A-Z]+(\.[A-Z0-9_]+)*|UTMA[A-Z]+(\.[A-Z0-9]+)*|UBLKCP(\.[A-Z]+)*|LDTM(\.[a-z0-9]+)*|SYNCS(\.[A-Z0-9]+)*|MUFU\.TANH|FFMA2|FMUL2|FADD2|REDG(\.[A-Za-z0-9]+)*|REDUX|WARPSYNC(\.[A-Z]+)*|FENCE\.VIEW\.ASYNC\.S)\b" \
    | sort | uniq -c | sort -rn | head -28
done
echo
echo "== kernels (demangled) and their resources"
for f in tc_gemm tc_bwd; do
  python3 - "$f.ptxas.log" <<'PY'
import re, subprocess, sys
text = open(sys.argv[1]).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(Used \d+ registers, used \d+ barriers)", text, flags=re.S):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if re.search(r"tc_gemm_kernel<(2, 0, 0, 2, false, false|1, 0, 0, 2, false, false|0, 0, 1, 2, false, false|0, 0, 0, 2, true, false)>|dw_xf_kernel<0, 1>|dx_red_kernel<0, true>", name):
        print(repr(name) + "\t" + m.group(2))
PY
done
