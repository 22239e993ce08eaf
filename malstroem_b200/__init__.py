"""malstroem_b200 — the raster hot path of SDFIdk/malstroem (fill, D8, accumulation, bluespot and watershed
labels, per-label reductions) as hand-written sm_100a CUDA behind the reference's numpy signatures.

    from malstroem_b200 import speedups; speedups.enable()      # plug into an installed malstroem
    from malstroem_b200.algorithms import fill, flow, label      # or call the mirrors directly
    from malstroem_b200.pipeline import RasterPipeline           # device-resident whole path (bench.py)
"""
__version__ = "0.1.0"
