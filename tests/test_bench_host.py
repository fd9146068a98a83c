"""Host-side pieces of bench.py that do not need a GPU: the clock/throttle summary."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location('skm_bench', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _row(sm, mx=1965, hw='Not Active', hwt='Not Active', swt='Not Active', cap='Not Active'):
    return f'{sm}, {mx}, 512.3, {hw}, {hwt}, {swt}, {cap}'


def test_clock_summary_uses_samples_inside_the_timed_window():
    b = _bench()
    rows = [(9.0, _row(300)), (10.01, _row(1900)), (10.05, _row(1920, cap='Active')),
            (10.09, _row(1910)), (11.0, _row(200, hw='Active'))]
    c = b.ClockSampler.summarize(rows, [10.0, 10.1])
    assert c['samples_in_timed_region'] == 3 and c['samples'] == 3
    assert c['sm_mhz'] == 1910.0 and c['sm_max_mhz'] == 1965.0
    assert c['reasons'] == ['sw_power_cap']       # the hw_slowdown sample lies outside the window


def test_clock_summary_falls_back_to_warmup_samples_and_skips_bad_lines():
    b = _bench()
    rows = [(9.9, _row(1800)), (9.95, 'garbage'), (9.97, '[N/A], 1965, 1, a, b, c, d'), (12.0, _row(100))]
    c = b.ClockSampler.summarize(rows, [10.0, 10.001])
    assert c['samples_in_timed_region'] == 0
    assert c['samples'] == 1 and c['sm_mhz'] == 1800.0
    c = b.ClockSampler.summarize([], [None, None])
    assert c['sm_mhz'] is None and c['samples'] == 0 and c['reasons'] == []
