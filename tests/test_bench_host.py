"""Host-side pieces of bench.py that run without a GPU: the clock sampler's fallbacks, the peak table, and the reference
arm's contract (`--impl reference` prints one JSON line carrying impl / cpu_baseline / e2e)."""
import json
import os
import subprocess
import sys
import time

from conftest import ROOT


def test_clock_sampler_degrades_without_a_gpu():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.02)
    out = s.stop()
    assert set(out) >= {'sm_mhz', 'sm_max_mhz', 'reasons', 'samples', 'source'}
    assert out['samples'] == 0 or out['sm_mhz'] > 0
    assert isinstance(out['reasons'], list)


def test_peaks_table_has_the_roofline_denominators():
    sys.path.insert(0, ROOT)
    import bench
    p = bench.load_peaks()
    assert p['hbm_gbs'] > 1000 and p['bf16_tflops_sustained'] > 100 and 'source' in p


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'nmpc_solves_per_s' and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['value'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
