"""`Modules` as the reference scripts and its pickles (`torch.save(model)` -> `Modules.Classifier`) know it:
a re-export of the B200-native operator surface, so `torch.load(model2load)` and `from Modules import *`
resolve to the CUDA-backed classes."""
from matcha_b200.hyper_sagnn import *  # noqa: F401,F403
from matcha_b200.hyper_sagnn import (Classifier, DataGenerator, EncoderLayer, FeedForward, MultiHeadAttention,  # noqa: F401
                                     MultipleEmbedding, PositionwiseFeedForward, SparseEmbedding, TiedAutoEncoder,
                                     activation, device, get_non_pad_mask)
