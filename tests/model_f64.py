"""Independent float64 numpy model of the RaBitQ estimator.

Not a restatement of the reference's instruction order: it re-derives the quantities from the definitions
(reference src/rabitq.rs:218-229 for the factors, :352-363 for the estimator) so it can catch a wrong
formula in BOTH the oracle and the CUDA path.  Agreement is to float32 round-off, not bit-exact.
"""
import numpy as np


def unpack_bits(words_u64: np.ndarray, dim: int) -> np.ndarray:
    """bit (i % 64) of word i // 64 -> array of 0/1 of length dim (src/utils.rs:53-61)."""
    w = np.asarray(words_u64, dtype=np.uint64)
    out = np.zeros(dim, np.int64)
    for i in range(dim):
        out[i] = (int(w[i // 64]) >> (i % 64)) & 1
    return out


def quantize_rne(residual: np.ndarray):
    """scalar_quantize on an AVX2 host (src/simd.rs:185-247): round-half-even of (r - lo) / delta, f32 arithmetic."""
    r = residual.astype(np.float32)
    lo, hi = np.float32(r.min()), np.float32(r.max())
    delta = np.float32((hi - lo) * np.float32(1.0 / 15.0))
    inv = np.float32(1.0) / delta
    q = np.rint(((r - lo) * inv).astype(np.float32)).astype(np.int64)
    return lo, delta, q


def planes_from_q(q: np.ndarray, dim: int) -> np.ndarray:
    """vector_binarize_query (src/simd.rs:83-107): plane b, word w, bit i%64 = bit b of q[i]."""
    W = dim // 64
    out = np.zeros(4 * W, np.uint64)
    for b in range(4):
        for i in range(dim):
            if (int(q[i]) >> b) & 1:
                out[b * W + i // 64] |= np.uint64(1) << np.uint64(i % 64)
    return out


def abdp(code_bits: np.ndarray, q: np.ndarray) -> int:
    """asymmetric_binary_dot_product == <x_bits, q_u> (src/utils.rs:113-135)."""
    return int((code_bits * q).sum())


def rough_f64(factor, ycd, lo, delta, sum_q, ab):
    ip, ppc, err, cds = [float(x) for x in factor]
    return cds + float(ycd) + float(lo) * ppc + (2.0 * ab - float(sum_q)) * ip * float(delta) - err * np.sqrt(float(ycd))


def rough_identity_f64(factor, ycd, lo, delta, q, code_bits):
    """rough = |x-c|^2 + |y-c|^2 + factor_ip * <s, lo + delta*q_u> - err*|y-c| with s = 2b-1."""
    ip, ppc, err, cds = [float(x) for x in factor]
    s = 2.0 * code_bits - 1.0
    rt = float(lo) + float(delta) * q.astype(np.float64)
    return cds + float(ycd) + ip * float((s * rt).sum()) - err * np.sqrt(float(ycd))
