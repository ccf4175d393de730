"""TEST INFRASTRUCTURE (oracle): restatement of the counter-based dropout masks the sm_100a kernels generate
(hybrid_vit_cascade_b200/csrc/hvc_common.cuh: mum32 / drop_rowkey / drop_colmul / drop_hash), so that train-mode parity can be
checked exactly: with the same mask, the CUDA path must match the reference arithmetic

    attn = dropout(softmax(q k^T * scale))          vit_components.py:47-49, :104-110
    x = dropout(proj(...)); mlp: Linear -> GELU -> Dropout -> Linear -> Dropout   vit_components.py:54-55,
                                                                                  hybrid_vit_backbone.py:75-81

The reference itself draws its masks from torch's Philox stream (nn.Dropout), which no fused kernel can reproduce
bit-for-bit; the mask generator is therefore part of the new design and parity under dropout is "same mask => same
numbers" plus the statistical checks in tests/ (keep rate, independence across sites/seeds, unbiasedness).
Only tests/ may import this module.
"""
import torch

_M32 = 0xFFFFFFFF


def _mum32(a, b):
    """a: int64 tensor holding uint32 values, b: python int (uint32). 32x32->64 multiply, fold hi ^ lo.
    Done in two 16-bit halves of b so the intermediate stays inside int64."""
    a = a & _M32
    lo = (a * (b & 0xFFFF))                    # < 2^48
    hi = (a * (b >> 16))                       # < 2^48 ; contributes << 16
    # full product = lo + (hi << 16): up to 2^64 -> split to avoid int64 overflow
    p_lo = (lo & _M32) + ((hi & 0xFFFF) << 16)             # < 2^33
    carry = p_lo >> 32
    p_hi = (lo >> 32) + (hi >> 16) + carry                 # high 32 bits
    return ((p_lo & _M32) ^ (p_hi & _M32)) & _M32


def _mul32(a, b):
    """low 32 bits of a * b; a, b: int64 tensors (or python ints) holding uint32 values"""
    a = a & _M32
    return ((a * (b & 0xFFFF)) + (((a * (b >> 16)) & 0xFFFF) << 16)) & _M32


def _mix32(c):
    """drop_mix32: 32-bit finaliser of (c + 1) * 0x9E3779B1"""
    x = _mul32((c + 1) & _M32, 0x9E3779B1)
    x = _mul32(x ^ (x >> 15), 0x85EBCA77)
    x = _mul32(x ^ (x >> 13), 0xC2B2AE3D)
    return x ^ (x >> 16)


def colmul(cols):
    """drop_colmul: blockmul(col / 128) * inblockmul(col % 128) mod 2^32, both factors odd"""
    cols = cols.to(torch.int64)
    return _mul32(_mix32((cols >> 7) ^ 0x5BD1E995) | 1, _mix32(cols & 127) | 1)


def rowkey(k0, k1, site, rows):
    """rows: int64 tensor of row ids -> odd uint32 keys (as int64)."""
    h = _mum32((rows & _M32) ^ k0, 0x9E3779B1)
    h = _mum32(h ^ site ^ k1, 0x85EBCA77)
    return _mum32((h + 0x6A09E667) & _M32, 0xC2B2AE3D) | 1


def keep_mask(seed_words, site, rows, cols, p):
    """Boolean keep mask [len(rows), len(cols)] for dropout probability p at `site`.
    seed_words: the two int32 words handed to the kernels (any signedness)."""
    k0, k1 = (int(w) & _M32 for w in seed_words)
    thr = min(int(p * 4294967296.0), _M32)
    rk = rowkey(k0, k1, int(site) & _M32, rows.to(torch.int64))
    h = _mul32(rk[:, None], colmul(cols)[None, :])             # drop_hash: the decision sits in the top bits of one 32-bit multiply
    return h >= thr


def inv_keep(p):
    thr = min(int(p * 4294967296.0), _M32)
    return 1.0 / (1.0 - thr / 4294967296.0)


class DropoutOracle:
    """Masks for the six dropout sites of every block of one backbone call (site = 8 * block + k,
    k: 0 self-attn probabilities, 1 self-attn proj, 2 cross-attn probabilities, 3 cross-attn proj, 4 MLP activation,
    5 MLP output), in the layouts the oracle functions use."""

    def __init__(self, seed_words, p, device="cpu"):
        self.seed = [int(w) for w in seed_words]
        self.p = float(p)
        self.device = device

    def attn(self, site, B, H, N, M):
        """(B, H, N, M) float mask * 1/(1-p): row id = (b*H + h)*N + q, col = key."""
        rows = torch.arange(B * H * N, device=self.device)
        cols = torch.arange(M, device=self.device)
        m = keep_mask(self.seed, site, rows, cols, self.p).view(B, H, N, M)
        return m.to(torch.float32) * inv_keep(self.p)

    def tokens(self, site, T, C):
        """(T, C) float mask * 1/(1-p): row = token index b*N + n, col = feature."""
        rows = torch.arange(T, device=self.device)
        cols = torch.arange(C, device=self.device)
        return keep_mask(self.seed, site, rows, cols, self.p).to(torch.float32) * inv_keep(self.p)


class TorchDropout:
    """nn.Dropout(p) in train mode exactly as the reference modules apply it (torch's own generator; vit_components.py:27-29,
    49, 55, 76-78, 110, 117; hybrid_vit_backbone.py:78, 80).  Used by the TIMED baselines (bench.py --impl reference / --impl eager),
    where the cost of the reference's dropout matters and the masks need not match the kernels'."""

    def __init__(self, p=0.1):
        self.p = float(p)

    def attn(self, site, B, H, N, M):
        import torch.nn.functional as F
        return lambda a: F.dropout(a, self.p, True)

    def apply_tokens(self, t):
        import torch.nn.functional as F
        return F.dropout(t, self.p, True)
