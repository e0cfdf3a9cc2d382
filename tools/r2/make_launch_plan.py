#!/usr/bin/env python
"""Launch plan of every conv call of one training step: the ABI call trace (bench.py --dump-trace) x snn_conv_plan (host-side
planning, no GPU needed).    python tools/r2/make_launch_plan.py profiles/r2/abi_trace_final.json > profiles/r2/launch_plan_final.md"""
import collections, json, math, re, sys
sys.path.insert(0, ".")
from snn_object_detectionddp_b200 import kernels as K

tr = json.load(open(sys.argv[1]))
rows = collections.OrderedDict()
for name, work in tr:
    if name not in ("snn_conv_fprop", "snn_conv_fprop_stats", "snn_conv_dgrad", "snn_conv_wgrad"):
        continue
    tag = work[2]
    m = re.match(r"g(\d) nb(\d+) (\d+)x(\d+) (\d+)(->|<-|x)(\d+)", tag)
    g, nb, h, w, cin, op, cout = int(m[1]), int(m[2]), int(m[3]), int(m[4]), int(m[5]), m[6], int(m[7])
    if (name, tag) in rows:
        rows[(name, tag)][0] += 1
        continue
    if op == "->":
        p = K.conv_plan("fprop", g, nb, h, w, cin, cout, out_f32=True, frames_per_step=64 if name.endswith("stats") else 0)
    elif op == "<-":
        # the ConvLSTM recurrent dgrad (W_h part: one frame, 1024 <- 4096) is the one dgrad with an fp32 output (dh accumulates over t)
        p = K.conv_plan("dgrad", g, nb, h, w, cin, cout, out_f32=(cout == 4 * cin and nb * 4 <= 256))
    else:
        p = K.conv_plan("wgrad", g, nb, h, w, cin, cout)
    rows[(name, tag)] = [1, p, work[1]]
print("# Launch plan of every conv call of one training step (configs[1]: T=4, B=64, 256x256), final code\n")
print("From `snn_conv_plan` (the host-side planning run without launching; `tests/test_plan.py` pins it on CPU) over the ABI call trace")
print("`abi_trace_final.json` (`tools/r2/make_launch_plan.py`).  box = images x rows x pixels of the staged pixel box; strip = row-strip mode")
print("((h, n, w) row order); fill = work items / (rounds x persistent CTAs or CTA pairs): what the static tile walk leaves idle in the last")
print("round.  dgrad outputs are bf16 except the ConvLSTM recurrent one (fp32, K split 9-fold); two-source (concat)")
print("convs are planned on their total Cin.\n")
print("| call | shape | launches | box | N tile | CTA pair | strip | stages x KB | K split | work items | CTAs (pairs) | rounds | fill | GFLOP |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = wasted = 0.0
for (name, tag), (n, p, fl) in rows.items():
    rounds = math.ceil(p["items"] / p["ctas"])
    fill = p["items"] / (rounds * (74 if p["pair"] else 148))
    tot += n * fl
    wasted += n * fl * (1 / fill - 1)
    print(f"| {name[9:]} | {tag} | {n} | {p['bn']}x{p['bh']}x{p['bw']} | {p['n_tile']} | {'yes' if p['pair'] else 'no'} | {'yes' if p['strip'] else 'no'} | "
          f"{p['stages']} x {p['stage_bytes'] // 1024} | {p['ksplit']} | {p['items']} | {p['ctas']} | {rounds} | {fill:.2f} | {fl / 1e9:.1f} |")
print(f"\nFLOP-weighted fill over all conv calls: {tot / (tot + wasted):.3f} (the rest is last-round idling of the static walk plus under-filled single rounds).")
