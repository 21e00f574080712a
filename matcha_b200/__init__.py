"""matcha_b200 -- B200-native (sm_100a) implementation of MATCHA's Hyper-SAGNN hyperedge-scoring hot path.

Public surface:
  matcha_b200.hyper_sagnn   mirror of the reference's Code/Modules.py classes (Classifier, MultipleEmbedding, ...)
  matcha_b200.sampler       device hash set of positive k-mers + GPU negative sampler
  matcha_b200.scorer        streaming all-pairs / k-way tuple scorers
  matcha_b200.trainer       fused, sync-free training step (optionally data-parallel over NCCL)
  matcha_b200.synthetic     seeded synthetic SPRITE-like inputs of the BASELINE.json shapes
The arithmetic lives in matcha_b200/csrc/*.cu behind the C ABI of include/matcha_b200.h.
"""
from ._lib import MatchaError, build, load  # noqa: F401

__version__ = "0.1.0"
