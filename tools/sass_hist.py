"""Static SASS opcode histogram + register/spill figures of one kernel of a model cubin.
usage: python tools/sass_hist.py <model> <kernel>"""
import collections, re, subprocess, sys
sys.path.insert(0, ".")
from triflow_b200 import workloads as W
from triflow_b200.model import Model
name, kern = sys.argv[1], sys.argv[2]
m = Model(**W.model_args(name), compiler="cuda")
cubin = m._cuda.variant(()).cubin_path
out = subprocess.run(["cuobjdump", "-sass", "-fun", kern, cubin], capture_output=True, text=True).stdout
ops = collections.Counter()
for line in out.splitlines():
    mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if mm:
        ops[mm.group(2).split(".")[0]] += 1
print(cubin, "total", sum(ops.values()))
print(" ".join("%s=%d" % kv for kv in ops.most_common(24)))
res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
lines = res.splitlines()
for i, l in enumerate(lines):
    if kern in l and i + 1 < len(lines):
        print(lines[i + 1].strip())
