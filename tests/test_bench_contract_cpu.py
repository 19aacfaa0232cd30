"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys.
Runs on 96 x 160 frames (bench.H / bench.W patched) so the CPU suite stays fast; the real arm uses 720p."""
import argparse
import json

import pytest

import bench

REQUIRED = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'impl', 'config', 'cpu_baseline', 'e2e')


@pytest.mark.parametrize('backbone', ['r50', 'swin_b'])
def test_reference_arm_json_line(monkeypatch, capsys, backbone):
    monkeypatch.setattr(bench, 'H', 96)
    monkeypatch.setattr(bench, 'W', 160)
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, backbone=backbone)
    bench.run_reference(args, rank=0)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['vs_baseline'] is None and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert ('Swin-B' in d['metric']) == (backbone == 'swin_b') and 'workload' in d['config']


def test_reference_arm_other_ranks_print_nothing(capsys):
    bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=0, backbone='r50'), rank=1)
    assert capsys.readouterr().out == ''
