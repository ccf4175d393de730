"""Statistics of the counter-based dropout mask the kernels regenerate (oracle/dropout_mask.py restates
hvc_common.cuh: drop_rowkey / drop_colmul / drop_hash).  CPU only: the GPU tests check that the kernels produce exactly this mask
(tests/test_dropout_gpu.py); here the mask itself is checked against Bernoulli(1 - p)."""
import torch

from oracle import dropout_mask as DM

SEED = [123456789, -987654321]


def _mask(p, rows=2048, cols=8192, site=3, seed=SEED):
    return DM.keep_mask(seed, site, torch.arange(rows), torch.arange(cols), p)


def test_keep_rate_and_binomial_spread():
    for p in (0.1, 0.5):
        m = _mask(p).float()
        n_r, n_c = m.shape
        assert abs(float(m.mean()) - (1 - p)) < 4 * (p * (1 - p) / m.numel()) ** 0.5 + 1e-4
        # per-row and per-column keep rates scatter like a binomial, i.e. rows/columns are neither biased nor too regular
        for dim, n in ((1, n_c), (0, n_r)):
            spread = float(m.mean(dim).std())
            expect = (p * (1 - p) / n) ** 0.5
            assert 0.85 * expect < spread < 1.15 * expect, (p, dim, spread, expect)


def test_no_correlation_between_neighbours_blocks_rows_sites_and_seeds():
    p = 0.1
    m = _mask(p).float() - (1 - p)
    noise = 5 * p * (1 - p) / (m.numel() ** 0.5)             # 5 sigma of the covariance estimate
    for lag in (1, 2, 3, 7, 32, 127, 128, 129, 256, 1024):  # 128 = the column-block size of the hash
        assert abs(float((m[:, :-lag] * m[:, lag:]).mean())) < noise, lag
    for lag in (1, 2, 64, 128, 1000):
        assert abs(float((m[:-lag] * m[lag:]).mean())) < noise, lag
    other_site = _mask(p, site=4).float() - (1 - p)
    other_seed = _mask(p, seed=[SEED[0] + 1, SEED[1]]).float() - (1 - p)
    assert abs(float((m * other_site).mean())) < noise and abs(float((m * other_seed).mean())) < noise


def test_gap_lengths_are_geometric():
    p = 0.1
    m = _mask(p, rows=64, cols=32768)
    gaps = torch.cat([(~r).nonzero().flatten().diff() for r in m]).float()
    assert abs(float(gaps.mean()) - 1 / p) < 0.15 and abs(float(gaps.std()) - ((1 - p) ** 0.5) / p) < 0.3
    # P(gap = 1) = p: adjacent drops happen as often as independence predicts
    assert abs(float((gaps == 1).float().mean()) - p) < 0.01


def test_inverse_keep_scale_is_unbiased():
    for p in (0.1, 0.25):
        m = _mask(p, rows=512, cols=4096).float()
        assert abs(float((m * DM.inv_keep(p)).mean()) - 1.0) < 5e-3


def test_every_column_pair_is_independent():
    """Round 2 mask: hash = rowkey * colmul(col), so two columns of one row are tied by a FIXED odd ratio colmul(c2) / colmul(c1).  For
    the constants in use the joint drop probability of EVERY pair among the first 384 columns (three 128-column blocks, 73 536 pairs) over
    2^18 rows sits within sampling noise of p^2, the per-row drop counts are binomial, and triples of neighbours co-drop at p^3."""
    p, rows, cols = 0.1, 1 << 18, 384
    drop = (~DM.keep_mask(SEED, 5, torch.arange(rows), torch.arange(cols), p)).float()
    joint = (drop.t() @ drop) / rows
    dev = joint - p * p
    dev.fill_diagonal_(0.0)
    sigma = (p * p * (1 - p * p) / rows) ** 0.5
    assert float(dev.abs().max()) < 5.5 * sigma, float(dev.abs().max()) / sigma       # max of 73 536 normal draws ~ 4.5 sigma
    cnt = drop.sum(1)
    assert abs(float(cnt.mean()) - cols * p) < 0.05 and abs(float(cnt.var()) / (cols * p * (1 - p)) - 1) < 0.02
    trip = (drop[:, :-2] * drop[:, 1:-1] * drop[:, 2:]).mean()
    assert abs(float(trip) - p ** 3) < 5e-5


def test_row_keys_are_odd_and_column_multipliers_distinct():
    rk = DM.rowkey(SEED[0] & 0xFFFFFFFF, SEED[1] & 0xFFFFFFFF, 9, torch.arange(100000))
    assert bool((rk & 1).all())
    cm = DM.colmul(torch.arange(32768))
    assert bool((cm & 1).all()) and cm.unique().numel() == 32768
