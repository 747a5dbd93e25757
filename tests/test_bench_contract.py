"""The JSON line `bench.py` prints is a contract with the driver.  This CPU test checks the shape of the last committed line
(profiles/r2_bench_final_default.json, written by `python bench.py` on a B200) and the pure helpers behind it."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    txt = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()
    return json.loads([l for l in txt if l.startswith("{")][-1])


def test_committed_bench_line_has_the_contract_keys():
    d = _line("r2_bench_final_default.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["dtype"] == "u8" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert isinstance(r["traffic"], int) and 0.5 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.5     # no wasted re-reads
    assert 0.3 < r["issue_bound"]["frac"] < 1.0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == bench.UNIT and e["h2d_bytes_per_step"] > 64 * 1241 * 376 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.02           # through host buffers is never faster than device-resident
    assert e["value"] == max(e[f]["value"] for f in ("per_keypoint_arrays", "packed_records", "map_associations")) and e["form"] in e
    assert d["gpu_launches"] > 0 and not d["clocks"]["reasons"]
    assert abs(d["value"] - d["config"]["features_per_step_per_gpu"] / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]


def test_e2e_block_picks_the_fastest_form_and_lists_all():
    forms = {"a": (10.0, 2.0, 100, "call a"), "b": (12.5, 1.6, 80, "call b"), "c": (11.0, 1.8, 60, "call c")}
    blk = bench.e2e_block(forms, d2h=7, call="x", extra={"k": 1})
    assert blk["form"] == "b" and blk["value"] == 12.5 and blk["h2d_bytes_per_step"] == 80 and blk["d2h_bytes_per_step"] == 7
    assert blk["ms_per_step"] == 1.6 and blk["k"] == 1 and blk["unit"] == bench.UNIT
    for name, (v, ms, h2d, c) in forms.items():
        assert blk[name] == {"value": v, "unit": bench.UNIT, "ms_per_step": ms, "h2d_bytes_per_step": h2d, "call": c}


def test_traffic_file_is_keyed_to_a_source_hash():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert len(t["_src_sha256"]) == 64 and set(t["_warp_instructions"]) >= {"pyramid", "fast", "blur", "describe", "search_frame"}
    # load_traffic() hands the numbers out only for the tree they were captured on
    got, why = bench.load_traffic()
    assert (got is not None) == (t["_src_sha256"] == bench.orb_source_hash()) and why
