"""A/B harness for kernel variants (run on the GPU box in ONE gpurun call).

Here (no GPU):   python -c "from granne_b200 import build; build.build_variant('w28', ['GB_MIN_BLOCKS=28', 'GB_STG_BYTES=2048'])"
On the box:      python tools/ab_bench.py granne_b200/libgranne_b200.so granne_b200/libgranne_b200_w28.so --config c2
                 (extra arguments go to bench.py; without --config it runs the 100M default, ~5 min per library)

For every library: the bit-exact parity subset (tests/test_parity_gpu.py + tests/test_golden.py) must pass, then bench.py
runs with that library (GRANNE_B200_LIB) and the JSON line is collected.  Prints one table; writes gpurun_out/ab_bench.jsonl.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(lib, steps, extra):
    env = dict(os.environ, GRANNE_B200_LIB=os.path.abspath(lib))
    t = subprocess.run([sys.executable, "-m", "pytest", "tests/test_parity_gpu.py", "tests/test_golden.py", "-m", "gpu", "-x",
                        "-q"], cwd=ROOT, env=env, capture_output=True, text=True)
    parity = t.returncode == 0
    line = None
    if parity:
        b = subprocess.run([sys.executable, "bench.py", "--steps", str(steps), "--cpu-seconds", "0"] + extra, cwd=ROOT, env=env,
                           capture_output=True, text=True)
        for out_line in b.stdout.splitlines():
            if out_line.startswith("{") and '"metric"' in out_line:
                line = json.loads(out_line)
    return parity, line, t.stdout[-400:]


def main():
    libs = [a for a in sys.argv[1:] if a.endswith(".so")]
    extra = [a for a in sys.argv[1:] if not a.endswith(".so")]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rows = []
    with open(os.path.join(ROOT, "gpurun_out", "ab_bench.jsonl"), "a") as f:
        for lib in libs:
            parity, line, tail = run(lib, 400, extra)
            rows.append((lib, parity, line))
            f.write(json.dumps({"lib": lib, "parity": parity, "bench": line, "pytest_tail": tail}) + "\n")
    print("%-44s %-7s %12s %12s %8s" % ("library", "parity", "value QPS", "e2e QPS", "roofline"))
    for lib, parity, line in rows:
        if line:
            print("%-44s %-7s %12.0f %12.0f %8.3f" % (os.path.basename(lib), parity, line["value"], line["e2e"]["value"],
                                                      line["roofline"]["frac"]))
        else:
            print("%-44s %-7s %12s" % (os.path.basename(lib), parity, "-"))


if __name__ == "__main__":
    main()
