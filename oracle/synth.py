"""Synthetic planted inputs for the hot path (SURVEY.md section 8d / BASELINE.md section 3).

Uniform random data degenerates to ~N singleton clusters, so the benchmark and the
parity tests use *planted* genomes: every contig belongs to one of ``G`` genomes
that share a TNF centroid and an abundance profile.
"""
from __future__ import annotations

import numpy as np


def make_contigs(n: int, nsamples: int, seed: int = 0, unique_lengths: bool = False, return_genome: bool = False):
    """Return ``(abundance[n, S], tnf[n, 103], lengths[n])`` float32/float32/int64.

    G = max(10, n // 50) genomes; TNF centroid c_g ~ N(0, I_103); abundance profile
    a_g ~ Gamma(1, 1)^S + 1e-3; contig i: g_i ~ U{0..G-1}, len_i = max(2000,
    floor(LogNormal(8.5, 1.0))), tnf_i = c_g + N(0, 0.3^2) * sqrt(2000 / len_i),
    abund_i = a_g * Gamma(20, 1/20)^S.
    """
    rng = np.random.default_rng(seed)
    g = max(10, n // 50)
    centroids = rng.standard_normal((g, 103), dtype=np.float32)
    profiles = (rng.gamma(1.0, 1.0, size=(g, nsamples)) + 1e-3).astype(np.float32)
    genome = rng.integers(0, g, size=n)
    lengths = np.maximum(2000, np.floor(rng.lognormal(8.5, 1.0, size=n))).astype(np.int64)
    if unique_lengths:
        # tie-free lengths make ``argsort(lengths)[::-1]`` independent of the NumPy build
        order = np.argsort(lengths, kind="stable")
        lengths[order] = lengths[order] + np.arange(n)  # strictly increasing along the sort
    noise = rng.standard_normal((n, 103), dtype=np.float32)
    scale = (0.3 * np.sqrt(2000.0 / lengths)).astype(np.float32)
    tnf = centroids[genome] + noise * scale[:, None]
    abundance = profiles[genome] * rng.gamma(20.0, 1.0 / 20.0, size=(n, nsamples)).astype(np.float32)
    out = (
        np.ascontiguousarray(abundance, dtype=np.float32),
        np.ascontiguousarray(tnf, dtype=np.float32),
        lengths,
    )
    return out + (genome,) if return_genome else out  # the planted genome of every contig (for ARI checks)


def make_latent(n: int, nlatent: int = 32, seed: int = 0, spread: float = 0.05,
                unique_lengths: bool = True):
    """Planted latent for clustering-only work: mu_g ~ N(0, I), x_i = mu_g + N(0, spread^2).

    Returns ``(latent[n, nlatent] float32, lengths[n] int64)``.
    """
    rng = np.random.default_rng(seed)
    g = max(10, n // 50)
    mu = rng.standard_normal((g, nlatent), dtype=np.float32)
    genome = rng.integers(0, g, size=n)
    latent = mu[genome] + np.float32(spread) * rng.standard_normal((n, nlatent), dtype=np.float32)
    lengths = np.maximum(2000, np.floor(rng.lognormal(8.5, 1.0, size=n))).astype(np.int64)
    if unique_lengths:
        order = np.argsort(lengths, kind="stable")
        lengths[order] = lengths[order] + np.arange(n)
    return np.ascontiguousarray(latent, dtype=np.float32), lengths
