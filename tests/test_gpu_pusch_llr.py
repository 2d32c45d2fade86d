"""GPU parity of the max-log LLR kernel against the CPU oracle (pinned to nr_ulsch_compute_llr)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Qm", [2, 4, 6, 8])
def test_pusch_llr_vs_oracle(ldpc, oracle, Qm):
    rng = np.random.default_rng(Qm)
    for n in (0, 1, 3, 8, 97, 3276, 3276 * 13):                  # empty, ragged tails, one symbol, one 273-PRB slot
        y = rng.integers(-32768, 32768, size=2 * n).astype(np.int16)
        y[:min(16, 2 * n)] = rng.choice(np.array([-32768, 32767, 0, -1], dtype=np.int16), size=min(16, 2 * n))
        mags = [rng.integers(0, 20000, size=2 * n).astype(np.int16) for _ in range(3)]
        got = ldpc.pusch_llr_host(Qm, y, *mags)
        if n:
            assert np.array_equal(got, oracle.ulsch_llr(Qm, y, *mags)), (Qm, n)
        else:
            assert got.size == 0
