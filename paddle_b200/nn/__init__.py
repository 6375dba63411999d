"""paddle.nn. Parity: python/paddle/nn/__init__.py."""
from . import functional, initializer, utils  # noqa: F401
from .layer import Layer, ParamAttr  # noqa: F401
from .container import LayerDict, LayerList, ParameterDict, ParameterList, Sequential  # noqa: F401
from .common import *  # noqa: F401,F403
from .common import (AlphaDropout, Bilinear, ChannelShuffle, CosineSimilarity, Dropout, Dropout2D, Dropout3D, Embedding,  # noqa: F401
                     FeatureAlphaDropout, Flatten, Fold, Identity, Linear, Pad1D, Pad2D, Pad3D, PairwiseDistance, PixelShuffle,
                     PixelUnshuffle, Unflatten, Unfold, Upsample, UpsamplingBilinear2D, UpsamplingNearest2D, ZeroPad1D, ZeroPad2D, ZeroPad3D)
from .conv_norm_pool import *  # noqa: F401,F403
from .conv_norm_pool import (AdaptiveAvgPool1D, AdaptiveAvgPool2D, AdaptiveAvgPool3D, AdaptiveMaxPool1D, AdaptiveMaxPool2D,  # noqa: F401
                             AdaptiveMaxPool3D, AvgPool1D, AvgPool2D, AvgPool3D, BatchNorm, BatchNorm1D, BatchNorm2D, BatchNorm3D,
                             Conv1D, Conv1DTranspose, Conv2D, Conv2DTranspose, Conv3D, Conv3DTranspose, FractionalMaxPool2D,
                             FractionalMaxPool3D, GroupNorm, InstanceNorm1D, InstanceNorm2D, InstanceNorm3D, LayerNorm, LocalResponseNorm,
                             LPPool1D, LPPool2D, MaxPool1D, MaxPool2D, MaxPool3D, MaxUnPool1D, MaxUnPool2D, MaxUnPool3D, RMSNorm,
                             SpectralNorm, SyncBatchNorm)
from .activation_loss import *  # noqa: F401,F403
from .activation_loss import (CELU, ELU, GELU, GLU, SELU, AdaptiveLogSoftmaxWithLoss, BCELoss, BCEWithLogitsLoss, CosineEmbeddingLoss,  # noqa: F401
                              CrossEntropyLoss, CTCLoss, GaussianNLLLoss, Hardshrink, Hardsigmoid, Hardswish, Hardtanh, HingeEmbeddingLoss,
                              HSigmoidLoss, HuberLoss, KLDivLoss, L1Loss, LeakyReLU, LogSigmoid, LogSoftmax, MarginRankingLoss, Maxout, Mish,
                              MSELoss, MultiLabelSoftMarginLoss, MultiMarginLoss, NLLLoss, PoissonNLLLoss, PReLU, ReLU, ReLU6, RNNTLoss, RReLU,
                              Sigmoid, Silu, SmoothL1Loss, SoftMarginLoss, Softmax, Softmax2D, Softplus, Softshrink, Softsign, Swish, Tanh,
                              Tanhshrink, ThresholdedReLU, TripletMarginLoss, TripletMarginWithDistanceLoss)
from .rnn import GRU, LSTM, RNN, BiRNN, GRUCell, LSTMCell, RNNCellBase, SimpleRNN, SimpleRNNCell  # noqa: F401
from .transformer import (MultiHeadAttention, Transformer, TransformerDecoder, TransformerDecoderLayer, TransformerEncoder,  # noqa: F401
                          TransformerEncoderLayer)
from .clip import ClipGradByGlobalNorm, ClipGradByNorm, ClipGradByValue  # noqa: F401
from . import decode, quant  # noqa: F401
from .decode import BeamSearchDecoder, dynamic_decode  # noqa: F401
