#!/bin/bash
# Full SASS listings of the hot kernels + a mnemonic summary of every kernel of the library -> profiles/
# usage: tools/sass_listing.sh r2
set -e
tag=${1:-r2}
cd "$(dirname "$0")/.."
lib=go-dsp_b200/lib/libgodsp_b200.so
mkdir -p profiles
cuobjdump -sass $lib > /tmp/all.sass
# kernel name fragments: <LA, LB, MODE, INV, PROF, TW2>; TW2 = 1: outer twiddle on the stores, 2: + stores into the peers' buffers, 3: aux product
for k in fft_tma14_kernelILi128ELi128ELi0ELb0ELb0ELi0E fft_tma14_kernelILi128ELi128ELi1ELb0ELb0ELi0E fft_tma14_kernelILi256ELi256ELi0ELb0ELb0ELi0E \
         fft_tma14_kernelILi256ELi256ELi1ELb0ELb0ELi0E fft_tma14_kernelILi256ELi128ELi0ELb0ELb0ELi0E fft_tma14_kernelILi1024ELi512ELi0ELb0ELb0ELi0E \
         fft_tma14_kernelILi256ELi256ELi1ELb0ELb0ELi1E fft_tma14_kernelILi256ELi256ELi1ELb0ELb0ELi2E fft_tma14_kernelILi256ELi256ELi0ELb0ELb0ELi3E \
         bluestein_small_kernelILi13ELi1E pwelch_bulk_kernel; do
  awk -v pat="$k" '/Function : /{f=index($0,pat)>0} f' /tmp/all.sass | gzip -9 > profiles/${tag}_sass_${k}.txt.gz
done
awk '/Function : /{f=index($0,"fft_tma_fused_kernelILb0ELb0")>0} f' /tmp/all.sass > profiles/${tag}_sass_fft_tma_fused_kernel_forward.txt
python3 - "$tag" <<'PY'
import re, sys, collections
tag = sys.argv[1]
out = ["SASS mnemonic counts (cuobjdump -sass go-dsp_b200/lib/libgodsp_b200.so, sm_100a), per kernel: static instruction counts.",
       "UTMALDG/UTMASTG = cp.async.bulk.tensor load/store (TMA), UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier ops,",
       "USETMAXREG = setmaxnreg, LDL/STL = spills. Full listings of the hot kernels: profiles/%s_sass_*.txt.gz" % tag, ""]
name, cnt, n, maxr = None, collections.Counter(), 0, 0
keep = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "USETMAXREG", "FENCE", "MEMBAR", "BAR", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "LDGSTS", "LDC", "REDG", "ATOMG", "LDL", "STL", "NANOSLEEP"]
def flush():
    if name and n:
        out.append(name); out.append("  instructions %d, highest register R%d" % (n, maxr))
        out.append("  " + "  ".join("%s=%d" % (k, cnt[k]) for k in keep if cnt[k])); out.append("")
for line in open("/tmp/all.sass"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); name, cnt, n, maxr = m.group(1), collections.Counter(), 0, 0; continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m:
        n += 1; cnt[m.group(1)] += 1
        for r in re.findall(r"\bR(\d+)\b", line): maxr = max(maxr, int(r))
flush()
open("profiles/%s_sass_summary.txt" % tag, "w").write("\n".join(out))
print("wrote profiles/%s_sass_summary.txt (%d kernels)" % (tag, sum(1 for l in out if l.startswith("_Z"))))
PY
