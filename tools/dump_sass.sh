#!/bin/bash
# Machine-code evidence for the shipped library: per kernel of libmerlin_b200.so (every cubin is sm_100a) the resource
# usage (cuobjdump -res-usage) and the histogram of the opcodes that carry the memory traffic -- global stores by
# width / cache hint (STG.E.EF.128 = the 16-byte streaming frame stores), bulk copies through the TMA unit
# (UBLKCP.G.S + UTMACMDFLUSH), shared-memory loads by width, global loads, atomics, votes / shuffles -- plus the total
# instruction count.  Runs on the build box (no GPU needed).
#     tools/dump_sass.sh > profiles/r02_sass_summary.txt
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
so=${1:-$root/ppo-2dgrid_b200/lib/libmerlin_b200.so}
echo "# SASS summary of $(basename "$so")  ($(date -u +%Y-%m-%dT%H:%MZ), $(/usr/local/cuda/bin/nvcc --version | grep release | sed 's/.*release //'))"
echo "# architectures: $(cuobjdump -lelf "$so" | sed 's/.*\.\(sm_[0-9a-z]*\)\.cubin/\1/' | sort | uniq -c | tr '\n' ' ')"
echo
echo "## resource usage (cuobjdump -res-usage)"
cuobjdump -res-usage "$so" 2>/dev/null | awk '/Function/{name=$2} /REG:/{print name, $0}' | while read -r name rest; do
  printf "%-88s %s\n" "$(echo "$name" | sed 's/:$//' | c++filt | cut -c1-86)" "$rest"
done
echo
echo "## opcode histogram per kernel (memory-traffic and warp-collective opcodes; total = all SASS instructions)"
cuobjdump -sass "$so" 2>/dev/null | python3 -c '
import re, sys, subprocess, collections
cur, stats = None, collections.OrderedDict()
pat = re.compile(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
keep = ("STG", "LDG", "LDS", "STS", "UBLKCP", "UTMA", "ATOM", "RED", "VOTE", "SHFL", "BAR", "BREV", "STL", "LDL", "FENCE", "MEMBAR", "CCTL", "SYNCS")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); stats[cur] = collections.Counter(); continue
    m = pat.match(line)
    if m and cur:
        op = m.group(1)
        stats[cur]["total"] += 1
        if op.startswith(keep):
            stats[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.split("\n")
for (k, c), name in zip(stats.items(), names):
    print(name[:120])
    print("    total %d | " % c.pop("total") + "  ".join("%s x%d" % kv for kv in sorted(c.items())))
'
