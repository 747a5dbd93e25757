"""A/B of the ORB bench stages across builds of the library: tools/ab_orb.py build/libcmos_X.so ..."""
import json
import os
import subprocess
import sys

for lib in sys.argv[1:]:
    env = dict(os.environ, CMOS_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "3", "--no-ba", "--no-cpu"], env=env,
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(lib, round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()})
    except Exception as e:  # noqa: BLE001
        print(lib, "failed", e, out.stderr[-300:])
