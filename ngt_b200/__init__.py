"""ngt_b200 -- B200 (sm_100a) engine for NGT's data-parallel hot path: graph beam search, exhaustive
kNN (linearSearch) and the kNN pass behind graph construction, behind NGT's own interfaces.

    ngt_b200.engine.GpuIndex   the device-resident index over the C ABI (include/ngtgpu.h)
    ngt_b200.synth             synthetic datasets of the BASELINE shapes

All compute is in ngt_b200/libngtgpu.so (hand-written CUDA); importing this package never imports oracle/.
"""
__all__ = ["engine", "synth"]
