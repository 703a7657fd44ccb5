"""CPU fuzz: the NumPy and the C restatement agree on random tiny cases with special floats (tests/fuzz_cases.py)."""
import numpy as np
import pytest

from fuzz_cases import make_case
from oracle import numpy_oracle as no, c_oracle as co


@pytest.mark.parametrize('block', range(6))
def test_numpy_and_c_oracles_agree_on_fuzz_cases(block):
    for seed in range(block * 40, block * 40 + 40):
        c = make_case(seed)
        N, h, w = c['shape']
        if h == 0 or w == 0:
            continue
        a = np.full((N, h, w), 4321, np.uint16)
        b = a.copy()
        pa = np.zeros((N, h, w, c['C']), np.float32)
        pb = pa.copy()
        no.eval_forest(c['forest'], c['depth'], a, c['r'], c['filt'], c['fclass'], c['scale'], probs_out=pa)
        co.eval_forest(c['forest'], c['depth'], b, c['r'], c['filt'], c['fclass'], c['scale'], probs_out=pb)
        assert np.array_equal(a, b), f'seed {seed}'
        assert np.array_equal(pa, pb, equal_nan=True), f'seed {seed}'
