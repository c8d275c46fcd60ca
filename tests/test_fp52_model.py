"""CPU check of the exact model behind experiments/dfma (FP64-pipe field / curve arithmetic, a round-2 candidate that the
product does not use yet): every floating-point step is exact, the column bookkeeping of fp52.cuh, the limb / magnitude
discipline of the G1 and G2 mixed additions, and the 8 x u32 conversions."""
import importlib.util
import os

import pytest

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'experiments', 'dfma', 'fp52_model.py')


@pytest.fixture(scope='module')
def model():
    spec = importlib.util.spec_from_file_location('fp52_model', _PATH)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_field_and_conversions(model):
    model._check()
    model._conv_check()


def test_mixed_additions_g1_g2(model):
    model.IMPL = model._Impl(False); model._ec_check()
    model.IMPL = model._Impl(True); model._ec_check()
    model._g2_check()
