"""bench.py pieces that need no GPU: the frozen flop formulas the roofline is computed from, the
clock sampler's handling of a timed window shorter than its sampling period, and the presence of the files the
roofline object reads its measured inputs from."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                             # noqa: E402


def test_frozen_flop_formula_of_the_survey():
    n, q = 1_000_000, 100
    # F(k) = N (Q (11 + 2 c_exp + 5 m_m) + (c_exp + 18) k + 75), m_m = 3 for k = 0, 4 otherwise; c_exp = 30 frozen
    assert bench.flops_per_eval(0, n, q) == n * (q * (11 + 60 + 15) + 75)
    assert bench.flops_per_eval(63, n, q) == n * (q * (11 + 60 + 20) + 48 * 63 + 75)
    step = sum(bench.flops_per_eval(k, n, q) for k in range(64))
    assert abs(step / 1e9 - 683.5) < 0.5                 # the 683.5 GF per step DESIGN.md quotes
    # the two other counts reported next to it are smaller (cheaper exp, executed flops only)
    assert bench.flops_executed(63, n, q) < bench.flops_per_eval(63, n, q, c_exp=14) < bench.flops_per_eval(63, n, q)
    assert bench.bytes_per_eval(63, n) == 8 * n * 64


def test_clock_sampler_falls_back_to_samples_around_a_short_window():
    row = lambda mhz, cap='Not Active': ['0', str(mhz), '1965', '400.0', 'Not Active', 'Not Active', 'Not Active', cap]
    s = bench.ClockSampler(0)
    now = time.perf_counter()
    s.t_on, s.t_off = now, now + 0.05
    s.nearby = [(now - 1.0, row(345)), (now - 0.2, row(1965)), (now + 0.1, row(1950, 'Active'))]
    out = s.summary()
    assert out['samples'] == 2 and out['samples_in_window'] == 0 and out['window_padding_s'] == 0.3
    assert out['sm_mhz'] in (1950.0, 1965.0) and out['reasons'] == ['sw_power_cap']     # the idle 345 MHz sample is out
    s.samples = [(now + 0.01, row(1965)), (now + 0.02, row(1960))]
    out = s.summary()
    assert out['samples'] == 2 and out['samples_in_window'] == 2 and out['window_padding_s'] == 0.0
    assert bench.ClockSampler(0).summary()['samples'] == 0


def test_measured_inputs_of_the_roofline_are_tracked():
    peaks = json.load(open(os.path.join(ROOT, 'profiles', 'fp64_peaks_r2.json')))
    assert 30.0 < peaks['dfma_tflops'] < 40.0 and 30.0 < peaks['dmma_m8n8k4_tflops'] < 40.0
    tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic_r2.json')))
    assert tr['launches'] == 64 and 1.0 <= tr['ratio_to_algorithmic'] < 1.2
    inv = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_inverse_traffic_r2.json')))
    assert inv['split']['ratio_to_algorithmic'] < 2.0 < inv['single_launch']['ratio_to_algorithmic']
