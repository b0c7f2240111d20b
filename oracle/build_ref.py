#!/usr/bin/env python3
"""Build oracle/_ref/libgsref.so from the reference's OWN shipped kernels.

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the product path.

What it does
------------
The reference (tatsuya-ogawa/GaussianSplattingMlx) ships its hot-path GPU kernels as
Metal-dialect source inside ``GaussianSplattingMlx/Slang/<name>_mlx.json`` (fields
``header`` + ``source``; loaded by ``Trainer/SlangKernelSpecLoader.swift:35-48``).  That
dialect is plain C++ once ``<metal_stdlib>`` is replaced by the small shim in
``oracle/metal_shim``.  This script reads those JSON files *where they lie* under
``/root/reference`` (never copied into the repository), wraps each kernel body into a C++
function ``body(thread_position_in_grid, buffers...)``, adds an OpenMP grid loop with a C ABI
(``ref_<kernel>(grid_x, grid_y, void** buffers)``) and compiles the lot with

    g++ -std=c++17 -O2 -fopenmp -fsingle-precision-constant -ffp-contract=off -fno-fast-math

into ``oracle/_ref/libgsref.so`` (git-ignored; travels to the GPU box through gpurun).  The
generated translation unit lives in a temporary directory and is deleted.

Two of the twelve kernels are cooperative (threadgroup barriers / simd_sum / float atomics)
and cannot run as independent per-thread bodies:

* ``radix_sort_tile_keys_fused_forward`` (slang/gaussian_tile_global_kernels.slang:143-305)
  is a *stable* LSD radix sort over (keyHigh[tileBits], keyLow[32]); it is replaced by
  ``std::stable_sort`` on the same key, which yields the identical permutation.
* ``gaussian_tile_global_backward`` (same file :648-881): its JSON ``header`` (the Slang
  autodiff-generated per-sample functions) IS compiled unchanged; the threadgroup choreography
  around them is restated as a serial per-pixel reverse loop (``ref_raster_backward`` below)
  that follows :696-723 and :756-847 statement by statement and accumulates the per-Gaussian
  sums in f64 (the reference's own sum order is nondeterministic: simd_sum + float atomics).
"""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("GSB_REFERENCE_ROOT", "/root/reference"))
REF_SLANG_DIR = REF_ROOT / "GaussianSplattingMlx" / "Slang"
REF_TRAINER_SWIFT = REF_ROOT / "GaussianSplattingMlx" / "Trainer" / "GaussianTrainer.swift"
OUT_DIR = HERE / "_ref"
OUT_SO = OUT_DIR / "libgsref.so"

# kernels that run as independent per-thread bodies
PER_THREAD = [
    "gaussian_projection_screen_fused_forward",
    "gaussian_projection_screen_fused_backward",
    "count_tiles_per_gaussian",
    "generate_keys",
    "compute_tile_ranges",
    "compute_tile_counts_from_ranges",
    "build_packed_tile_indices",
    "gaussian_tile_global_forward",
    "ssim_forward",
    "ssim_backward",
]

CXXFLAGS = [
    "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared",
    "-fsingle-precision-constant", "-ffp-contract=off", "-fno-fast-math",
    "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-narrowing",
]


def _strip_includes(header: str) -> str:
    out = []
    for line in header.splitlines():
        s = line.strip()
        if s.startswith("#include") or s == "using namespace metal;":
            continue
        out.append(line)
    return "\n".join(out)


def _ctype(metal_type: str) -> str:
    # e.g. "float device*" -> "float*"
    return metal_type.replace("device", "").replace("  ", " ").strip()


def _emit_per_thread(name: str, spec: dict) -> str:
    params = spec["buffer_parameters"]
    sig = ", ".join(f"{_ctype(p['type'])} {p['name']}" for p in params)
    call = ", ".join(f"({_ctype(p['type'])})b[{i}]" for i, p in enumerate(params))
    return f"""
namespace k_{name} {{
{_strip_includes(spec['header'])}
static inline void body(uint3 thread_position_in_grid, {sig})
{{
{spec['source']}
}}
}}  // namespace
extern "C" void ref_{name}(unsigned gx, unsigned gy, void** b)
{{
    const long total = (long)gx * (long)gy;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < total; ++i)
        k_{name}::body(uint3((unsigned)(i % gx), (unsigned)(i / gx), 0u), {call});
}}
"""


RASTER_BWD_DRIVER = r"""
// Serial restatement of the threadgroup choreography of gaussian_tile_global_backward
// (slang/gaussian_tile_global_kernels.slang:648-881) around the UNCHANGED generated
// per-sample functions of the JSON header.  One pixel = one serial reverse loop.
extern "C" void ref_raster_backward(
    const float* packed, const int* packedTileIndices, const unsigned* tileCounts,
    const float* cotColor, const float* cotDepth, const float* cotAlpha,
    const float* outColor, const float* outDepth, const unsigned* renderCounts,
    const float* outAlpha, const unsigned* lastContrib, double* gradPacked64)
{
    using namespace k_gaussian_tile_global_backward;
    const unsigned maxTilePairs = renderCounts[1], gridW = renderCounts[2];
    const unsigned tileW = renderCounts[3], tileH = renderCounts[4];
    const unsigned imageW = renderCounts[5], imageH = renderCounts[6], whiteBg = renderCounts[7];
    const long P = (long)imageW * imageH;
    #pragma omp parallel for schedule(dynamic, 64)
    for (long p = 0; p < P; ++p) {
        const unsigned pixelY = (unsigned)(p / imageW), pixelX = (unsigned)(p % imageW);
        const unsigned tileIndex = (pixelY / tileH) * gridW + (pixelX / tileW);
        const unsigned count = tileCounts[tileIndex];
        const size_t tileBase = (size_t)tileIndex * maxTilePairs;
        // :696-723
        const float cx = cotColor[p * 3], cy = cotColor[p * 3 + 1], cz = cotColor[p * 3 + 2];
        const float trans = 1.0f - outAlpha[p];
        const float bg = whiteBg != 0 ? trans : 0.0f;
        TileGlobalPixelState_0 cur, cot;
        cur.colorX_2 = outColor[p * 3 + 0] - bg;
        cur.colorY_2 = outColor[p * 3 + 1] - bg;
        cur.colorZ_2 = outColor[p * 3 + 2] - bg;
        cur.depth_2 = outDepth[p];
        cur.trans_0 = trans;
        cot.colorX_2 = cx; cot.colorY_2 = cy; cot.colorZ_2 = cz;
        cot.depth_2 = cotDepth[p];
        cot.trans_0 = -cotAlpha[p] + (whiteBg != 0 ? (cx + cy + cz) : 0.0f);
        const float px = (float)pixelX, py = (float)pixelY;
        const unsigned nContrib = lastContrib[p];
        // :756-847 (the chunking only stages data; order of ii is count-1 ... 0)
        for (long ii = (long)count - 1; ii >= 0; --ii) {
            if ((unsigned)ii >= nContrib) continue;
            const unsigned gi = (unsigned)packedTileIndices[tileBase + ii];
            const float* g = packed + (size_t)gi * 11;
            const float meanX = g[0], meanY = g[1], c00 = g[2], c01 = g[3], c10 = g[4], c11 = g[5];
            const float colorX = g[6], colorY = g[7], colorZ = g[8], opacity = g[9], depth = g[10];
            TileGlobalSample_0 sample = evaluateTileGlobalSample_0(
                meanX, meanY, c00, c01, c10, c11, opacity, colorX, colorY, colorZ, depth, px, py);
            TileGlobalPixelState_0 prev = undoTileGlobalPixelState_0(&cur, &sample);
            DiffPair_TileGlobalPixelState_0 dpPrev;
            dpPrev.primal_0 = prev;
            dpPrev.differential_0 = TileGlobalPixelState_x24_syn_dzero_0();
            DiffPair_TileGlobalSample_0 dpSample;
            dpSample.primal_0 = sample;
            dpSample.differential_0 = TileGlobalSample_x24_syn_dzero_0();
            s_bwd_updateTileGlobalPixelState_0(&dpPrev, &dpSample, &cot);
            cur = prev;
            cot = dpPrev.differential_0;
            DiffPair_float_0 d[11];
            const float prim[11] = {meanX, meanY, c00, c01, c10, c11, opacity, colorX, colorY, colorZ, depth};
            for (int k = 0; k < 11; ++k) { d[k].primal_0 = prim[k]; d[k].differential_0 = 0.0f; }
            TileGlobalSample_0 dS = dpSample.differential_0;
            s_bwd_evaluateTileGlobalSample_0(&d[0], &d[1], &d[2], &d[3], &d[4], &d[5], &d[6],
                                             &d[7], &d[8], &d[9], &d[10], px, py, &dS);
            // packed column order: mean(2) conic(4) color(3) opacity depth
            const int col[11] = {0, 1, 2, 3, 4, 5, 9, 6, 7, 8, 10};
            double* out = gradPacked64 + (size_t)gi * 11;
            for (int k = 0; k < 11; ++k) {
                const double v = (double)d[k].differential_0;
                #pragma omp atomic
                out[col[k]] += v;
            }
        }
    }
}
"""

STABLE_SORT = r"""
#include <vector>
#include <numeric>
// Stand-in for radix_sort_tile_keys_fused_forward (slang/gaussian_tile_global_kernels.slang:143-305):
// a stable ascending sort on (keyHigh & ((1<<highBitsRoundedUpTo4)-1), keyLow).  The reference sorts
// ceil(tileBits/4) 4-bit digits of keyHigh, i.e. the low 4*ceil(tileBits/4) bits.
extern "C" void ref_stable_sort_tile_keys(
    const unsigned* keysHigh, const unsigned* keysLow, const unsigned* values, unsigned M, unsigned tileBits,
    unsigned* sortedHigh, unsigned* sortedLow, unsigned* sortedValues)
{
    unsigned passes = (tileBits + 3u) / 4u; if (passes < 1u) passes = 1u;
    const unsigned bits = passes * 4u;
    const unsigned mask = bits >= 32u ? 0xffffffffu : ((1u << bits) - 1u);
    std::vector<unsigned> perm(M);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](unsigned a, unsigned b) {
        const unsigned ha = keysHigh[a] & mask, hb = keysHigh[b] & mask;
        if (ha != hb) return ha < hb;
        return keysLow[a] < keysLow[b];
    });
    for (unsigned i = 0; i < M; ++i) {
        sortedHigh[i] = keysHigh[perm[i]]; sortedLow[i] = keysLow[perm[i]]; sortedValues[i] = values[perm[i]];
    }
}
"""


# ------------------------------------------------------------------------------------------------
# The three densification kernels are NOT shipped as JSON: they are inline Metal strings handed to
# MLXFast.metalKernel(name:inputNames:outputNames:source:) in Trainer/GaussianTrainer.swift:321-427.
# Their bodies are extracted from the Swift file where it lies and wrapped with the buffer signature MLX
# generates for them (pointer per array input/output, `<name>_shape` int arrays, scalars by reference).
# ------------------------------------------------------------------------------------------------
INLINE_KERNELS = {
    # name: (signature, call expression) — b[] is the void* buffer table handed to ref_<name>()
    "accum_grad_norm": (
        "const float* xyz_grad, const float* accum_in, const int* accum_in_shape, float* accum_out",
        "(const float*)b[0], (const float*)b[1], (const int*)b[2], (float*)b[3]"),
    "classify_gaussians": (
        "const float* grad_accum, const int* grad_accum_shape, const float* denom_accum, const float* scales, "
        "const int* scales_shape, const float* opacity, const float& grad_threshold, const float& max_scale_thresh, "
        "const float& min_opacity_thresh, const int& allow_densify, int* actions, int* output_counts",
        "(const float*)b[0], (const int*)b[1], (const float*)b[2], (const float*)b[3], (const int*)b[4], (const float*)b[5], "
        "*(const float*)b[6], *(const float*)b[7], *(const float*)b[8], *(const int*)b[9], (int*)b[10], (int*)b[11]"),
    "build_densify_output_map": (
        "const int* actions, const int* actions_shape, const int* offsets, int* gather_indices, int* noise_mode",
        "(const int*)b[0], (const int*)b[1], (const int*)b[2], (int*)b[3], (int*)b[4]"),
}


def _inline_kernel_source(swift: str, name: str) -> str:
    m = re.search(r'name:\s*"' + re.escape(name) + r'".*?source:\s*"""(.*?)"""', swift, flags=re.S)
    if not m:
        raise RuntimeError(f"inline Metal kernel {name} not found in {REF_TRAINER_SWIFT}")
    return m.group(1)


def _emit_inline(name: str, body: str) -> str:
    sig, call = INLINE_KERNELS[name]
    return f"""
namespace k_{name} {{
static inline float fmax(float a, float b) {{ return a > b ? a : b; }}
static inline void body(uint3 thread_position_in_grid, {sig})
{{
{body}
}}
}}  // namespace
extern "C" void ref_{name}(unsigned gx, unsigned gy, void** b)
{{
    const long total = (long)gx * (long)gy;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < total; ++i)
        k_{name}::body(uint3((unsigned)(i % gx), (unsigned)(i / gx), 0u), {call});
}}
"""


def generate_source() -> str:
    parts = [
        "// GENERATED in a temporary directory by oracle/build_ref.py — never committed.\n",
        "#include <metal_stdlib>\n#include <cstddef>\nusing namespace metal;\n",
    ]
    for name in PER_THREAD:
        spec = json.loads((REF_SLANG_DIR / f"{name}_mlx.json").read_text())
        parts.append(_emit_per_thread(name, spec))
    bwd = json.loads((REF_SLANG_DIR / "gaussian_tile_global_backward_mlx.json").read_text())
    parts.append("namespace k_gaussian_tile_global_backward {\n" + _strip_includes(bwd["header"]) + "\n}\n")
    parts.append(RASTER_BWD_DRIVER)
    parts.append(STABLE_SORT)
    swift = REF_TRAINER_SWIFT.read_text()
    for name in INLINE_KERNELS:
        parts.append(_emit_inline(name, _inline_kernel_source(swift, name)))
    parts.append('extern "C" int ref_abi_version() { return 2; }\n')
    return "".join(parts)


def build(verbose: bool = True) -> Path | None:
    """Returns the path of the built library, or None when /root/reference is absent."""
    if not REF_SLANG_DIR.is_dir():
        if verbose:
            print(f"[build_ref] {REF_SLANG_DIR} not present - keeping prebuilt {OUT_SO} (exists={OUT_SO.exists()})")
        return OUT_SO if OUT_SO.exists() else None
    OUT_DIR.mkdir(exist_ok=True)
    tmp = Path(tempfile.mkdtemp(prefix="gsref_"))
    try:
        src = tmp / "gsref_generated.cpp"
        src.write_text(generate_source())
        # empty companions of <metal_stdlib>
        for extra in ("metal_math", "metal_texture"):
            (tmp / extra).write_text("#pragma once\n")
        cmd = ["g++", *CXXFLAGS, "-isystem", str(HERE / "metal_shim"), "-isystem", str(tmp),
               str(src), "-o", str(OUT_SO)]
        if verbose:
            print("[build_ref]", " ".join(cmd))
        subprocess.run(cmd, check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT_SO


if __name__ == "__main__":
    p = build()
    print(p)
    sys.exit(0 if p else 1)
