"""CPU tier: the roofline byte model of bench.py (algorithmic HBM bytes per input sample of a K2 step) against the
figures DESIGN.md derives by hand for the plans the operator reports."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    spec = importlib.util.spec_from_file_location("bench_for_model", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.fir_bytes_model


def plan(B, P, **kw):
    p = {"levels": [{"B": B, "P": P}], "tail_pf": 2, "t_batch": 4, "t_far": 0, "far_e": 0, "stagger": 1, "merge": 0, "pipe": 0, "pipe_pf": 0}
    p.update(kw)
    return p


def test_headline_single_tier(model):
    total, parts = model(plan(4096, 32), 256, 4096, 1)
    # fused kernel 112, per-block MAC 4 x 32 + V + Y = 160, batch tier (26 H rows + 29 X rows + 4 V rows) x 16 / 4 = 236
    assert parts["fir_level0"] == 112 and parts["fir_mac"] == 160 and parts["fir_mac_batch"] == 236
    assert total == 508


def test_one_grid_tail_is_the_same_bytes(model):
    total, parts = model(plan(4096, 32, merge=1), 256, 4096, 1)
    assert total == 508 and parts["fir_tail"] == 396 and "fir_mac" not in parts


def test_two_tiers(model):
    # far tier of 12 one near period ahead: near tier partitions 6..17, far 18..31
    total, parts = model(plan(4096, 32, t_far=12, far_e=4, stagger=0), 256, 4096, 1)
    assert parts["fir_mac_batch"] == 16 * ((12 + 3) + 12 + 4 + 4) / 4
    assert parts["fir_mac_batch_far"] == 16 * ((14 + 11) + 14 + 12) / 12
    assert abs(total - (112 + 160 + 140 + 68)) < 1e-9
    # 2048-frame blocks: 64 partitions, far tier of 8 one near period ahead (the planner's default there)
    total, _ = model(plan(2048, 64, t_far=8, far_e=4, stagger=0), 256, 2048, 1)
    assert total == 610
    # without the look-ahead (DESIGN.md's 594) and a single tier (764)
    assert model(plan(2048, 64, t_far=8, far_e=0, stagger=0), 256, 2048, 1)[0] == 594
    assert model(plan(2048, 64), 256, 2048, 1)[0] == 764


def test_shared_filter_rows_are_not_counted(model):
    # one IR for all channels: its rows stay in L2 (h = 0)
    total, parts = model(plan(4096, 16), 256, 4096, 0)
    assert parts["fir_level0"] == 112 - 32 and parts["fir_mac"] == 16 * (4 + 2)
    assert parts["fir_mac_batch"] == 16 * ((10 + 3) + 4) / 4


def test_untailed_and_multilevel_plans(model):
    # 13000 taps: 4 partitions, no batch: MAC over partitions 2, 3
    total, parts = model({"levels": [{"B": 4096, "P": 4}], "tail_pf": 2, "t_batch": 0, "pipe": 0}, 3, 4096, 1)
    assert parts["fir_mac"] == 16 * (2 * 2 + 1) and parts["fir_mac_batch"] == 0
    # two levels (2048 + 4096, tail on the upper one): staging kernels are part of the step
    total, parts = model({"levels": [{"B": 2048, "P": 2}, {"B": 4096, "P": 31}], "tail_pf": 1, "t_batch": 4, "pipe": 0}, 256, 2048, 1)
    assert parts["stash_unstash"] > 0 and total == pytest.approx(612, abs=1)
