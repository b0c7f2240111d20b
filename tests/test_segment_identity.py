"""The identity the segmented raster backward (raster.cu, RasterCkpt) rests on, in float64 numpy.

The reference walks a pixel's list back to front (slang/gaussian_tile_global_kernels.slang:648-881): with T the
transmittance AFTER Gaussian i and kT the accumulated "colour behind" term,
    prevT = T / (1 - a_i);  d = cot . c_i - kT;  g_alpha_i = prevT * d;  kT <- kT + a_i * d;  T <- prevT.
A segment [s, e) of the list can be differentiated on its own when it starts from the forward's state at e:
    T(e)  = prod_{j < e} (1 - a_j)                                   (the checkpointed transmittance)
    kT(e) = (cot . sum_{j >= e} w_j c_j + T_final * kT_init) / T(e)   (w_j = T(j) a_j: the later segments' colour sums)
This test checks that closed form against the recursion, including a pixel that terminates inside a segment.
"""
import numpy as np


def forward(alpha, color, seg):
    """Blend front to back; returns final T, final colour, per-segment colour sums and T at the segment ends."""
    n = len(alpha)
    T, sums, t_end = 1.0, [], []
    acc = np.zeros(3)
    for i in range(n):
        acc = acc + T * alpha[i] * color[i]
        T = T * (1.0 - alpha[i])
        if (i + 1) % seg == 0 or i + 1 == n:
            sums.append(acc.copy()); t_end.append(T)
            acc = np.zeros(3)
    return T, np.sum(sums, axis=0), sums, t_end


def backward_range(alpha, color, cot, lo, hi, T_after, kT_after):
    """The reference's back-to-front recursion over Gaussians [lo, hi), started from (T, kT) after Gaussian hi - 1."""
    T, kT = T_after, kT_after
    g_alpha = np.zeros(hi - lo)
    g_color = np.zeros((hi - lo, 3))
    for i in range(hi - 1, lo - 1, -1):
        prevT = T / (1.0 - alpha[i])
        d = cot @ color[i] - kT
        g_alpha[i - lo] = prevT * d
        g_color[i - lo] = prevT * alpha[i] * cot
        kT = kT + alpha[i] * d
        T = prevT
    return g_alpha, g_color


def test_segment_start_state_reproduces_the_whole_list_recursion():
    rng = np.random.default_rng(0)
    n, seg = 1000, 256
    alpha = rng.uniform(0.0, 0.05, n)
    alpha[rng.integers(0, n, 20)] = 0.99                     # clamped samples
    color = rng.uniform(0.0, 1.5, (n, 3))
    cot = rng.standard_normal(3)
    kT_init = 0.37                                            # -cotAlpha + white-background term
    for last in (n, 700, 512, 300):                           # lastContrib: end of list, inside / at the end of a segment
        a, c = alpha[:last], color[:last]
        T_fin, _, sums, t_end = forward(a, c, seg)
        ga_ref, gc_ref = backward_range(a, c, cot, 0, last, T_fin, kT_init)
        nseg = len(sums)
        for s in range(nseg):
            lo, hi = s * seg, min((s + 1) * seg, last)
            if s == nseg - 1:                                 # the block's final item starts from the final state
                T_e, kT_e = T_fin, kT_init
            else:                                             # an interior item starts from the forward's checkpoint
                T_e = t_end[s]
                kT_e = (cot @ np.sum(sums[s + 1:], axis=0) + T_fin * kT_init) / T_e
            ga, gc = backward_range(a, c, cot, lo, hi, T_e, kT_e)
            assert np.allclose(ga, ga_ref[lo:hi], rtol=1e-9, atol=1e-12)
            assert np.allclose(gc, gc_ref[lo:hi], rtol=1e-9, atol=1e-12)


def test_reference_rounding_of_the_final_transmittance_scales_the_whole_chain():
    """The reference starts from T = 1 - out_alpha in f32; the kernel applies the same factor to the checkpoint T so that a
    segment sees the transmittances the whole-list chain would have produced."""
    T_fin = np.float32(3.1234567e-5)
    out_alpha = np.float32(1.0) - T_fin
    T_ref = np.float32(1.0) - out_alpha                       # what the reference's chain starts from
    rho = float(T_ref) / float(T_fin)
    assert abs(rho - 1.0) > 1e-5                              # the rounding is visible ...
    alpha = np.array([0.1, 0.2, 0.3])
    T_e_true = float(T_fin) / np.prod(1.0 - alpha)            # transmittance three Gaussians earlier
    T_e_chain = float(T_ref) / np.prod(1.0 - alpha)           # what the reference's division chain yields there
    assert np.isclose(T_e_true * rho, T_e_chain, rtol=1e-12)  # ... and is exactly the factor applied to the checkpoint
