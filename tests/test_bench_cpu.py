"""bench.py on a machine without a GPU: the `--impl reference` arm (the reference's CPU
implementation of the path through the checker, what the driver times beside the B200 arm) runs
here at a reduced size and prints the contract's JSON line; the B200 arm refuses to run without a
CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=300):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True,
                          text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--width", "192", "--height", "128",
                  "--spp", "1", "--cpu-seconds", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_does_not_map_the_product_library():
    """The driver records which .so files each arm loaded: the reference arm is the CPU checker alone and
    must not even map libspb200.so ("it dirties the record")."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--width', '96', '--height', '64', "
            "'--spp', '1']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "sys.stderr.write('MAPPED_PRODUCT=%%d MAPPED_CHECKER=%%d\\n' %% ('libspb200' in maps, 'libsporacle' in maps or 'libspref' in maps))\n"
            % os.path.join(ROOT, "bench.py"))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "MAPPED_PRODUCT=0 MAPPED_CHECKER=1" in p.stderr, p.stderr[-500:]


def test_b200_arm_refuses_to_run_without_a_gpu():
    p = run_bench("--steps", "1", "--warmup", "1", "--width", "64", "--height", "48", "--spp", "1", "--quick",
                  timeout=120)
    assert p.returncode != 0
    assert not any(l.startswith("{") for l in p.stdout.splitlines())
