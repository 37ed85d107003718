"""B200-native drop-in for the aggregate_2p5d hot path of VisSatSatelliteStereo.

Module layout mirrors the reference's (aggregate_2p5d, aggregate_2p5d_util, produce_dsm,
coordinate_system, lib.*, colmap.read_dense); all compute goes through the sm_100a kernels in
``csrc/`` via the C-ABI declared in ``include/vissat_b200.h`` (ctypes loader: ``_native``).
There is no CPU fallback: without the built library or a CUDA device the compute entry points raise.
"""
__version__ = '0.1.0'
