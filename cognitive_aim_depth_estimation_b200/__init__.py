"""B200-native forward path of the Cognitive-Aim depth model.

Public surface mirrors the reference's `src/model.py`: `create_model`, `CognitiveAimModel.forward`,
`CognitiveAimModel.forward_with_guidance`.  Compute runs in hand-written sm_100a kernels inside
`libcogaim_b200.so`; there is no CPU fallback.
"""
__version__ = "0.1.0"
