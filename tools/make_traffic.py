"""profiles/traffic.json from an ncu --set full summary (tools/ncu_summary.py output) of ONE 64-frame step:
per bench stage, DRAM bytes read + written per launch (summed over the stage's kernels).

  python tools/make_traffic.py <summary.json> profiles/traffic.json [<sha256 file written on the box at capture time>]

`_src_sha256` is bench.orb_source_hash() of the tree the capture ran on; bench.py reports `roofline.traffic` only when it
matches the tree being benchmarked."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

STAGE = {"k_level0": "pyramid", "k_resize": "pyramid", "k_resize2": "pyramid", "k_borders": "pyramid", "k_fast": "fast", "k_octree": "quadtree", "k_blur": "blur",
         "k_describe": "describe", "k_build_grid": "grid", "k_sf_lists": "search_frame", "k_sf_replay": "search_frame"}
d = json.load(open(sys.argv[1]))
out, dur, winst = {}, {}, {}
# the capture may span several steps: keep ONE whole step = the launches from one k_level0 up to the next
kn = lambda r: r["kernel"].split("::")[-1].split("<")[0].replace("void ", "").strip()
starts = [i for i, r in enumerate(d["launches"]) if kn(r) == "k_level0"]
launches = d["launches"][starts[-2]:starts[-1]] if len(starts) >= 2 else d["launches"]
out["_launches_in_step"] = len(launches)
for r in launches:
    name = r["kernel"].split("::")[-1].split("<")[0].replace("void ", "").strip()
    st = STAGE.get(name)
    if st is None or "dram_traffic" not in r:
        continue
    out[st] = out.get(st, 0) + int(r["dram_traffic"])
    dur[st] = dur.get(st, 0.0) + r.get("duration_us", 0.0)
    winst[st] = winst.get(st, 0) + int(r.get("warp_instructions", 0))
if len(sys.argv) > 3:
    out["_src_sha256"] = open(sys.argv[3]).read().strip()
else:
    import bench
    out["_src_sha256"] = bench.orb_source_hash()
out["_source"] = f"{d['report']}: dram__bytes_read.sum + dram__bytes_write.sum per launch, one 64-frame step"
out["_ncu_duration_us"] = {k: round(v, 1) for k, v in dur.items()}
out["_warp_instructions"] = winst      # smsp__inst_executed.sum per stage of the same step (the path is instruction-issue bound)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
