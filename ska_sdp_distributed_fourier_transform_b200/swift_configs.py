"""
Named SwiFTly parameter sets.

The reference ships a catalogue of 244 entries (``swift_configs.py``); entries are plain
``SwiftlyConfig`` keyword dictionaries.  This module carries the parameter sets the
BASELINE benchmark is quoted on (none of which is in the reference catalogue, see
SURVEY.md section 8d) plus the reference's unit-test set; any dictionary with the same keys --
including every entry of the reference catalogue whose FFT lengths are powers of two --
can be passed to ``SwiftlyConfig(**params)``.
"""

SWIFT_CONFIGS = {
    # reference tests/test_core.py:20-27, tests/test_api.py:32-40
    "1k[1]-n512-256": dict(W=13.5625, fov=1.0, N=1024, yB_size=416, yN_size=512,
                           xA_size=228, xM_size=256),
    # BASELINE configs[1..3]
    "8k[1]-n4k-2k": dict(W=13.5625, fov=1.0, N=8192, yB_size=2048, yN_size=4096,
                         xA_size=1024, xM_size=2048),
    "32k[1]-n8k-4k": dict(W=13.5625, fov=1.0, N=32768, yB_size=4096, yN_size=8192,
                          xA_size=2048, xM_size=4096),
    "64k[1]-n16k-4k": dict(W=13.5625, fov=1.0, N=65536, yB_size=8192, yN_size=16384,
                           xA_size=2048, xM_size=4096),
}
