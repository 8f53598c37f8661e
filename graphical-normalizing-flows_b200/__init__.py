"""gnf_b200 — B200-native (sm_100a) hot path of Graphical Normalizing Flows.

Drop-in for the reference's ``models`` package on the density-evaluation / training path:
same class names, constructor signatures, attributes and state_dict keys; every forward /
backward runs in hand-written CUDA kernels behind the C-ABI of ``include/gnf.h``
(``libgnf_sm100.so``).  There is no CPU or eager-PyTorch fallback.
"""
from . import _lib, ops
from .conditioners import (AutoregressiveConditioner, Conditioner, ConditionnalMADE, CouplingConditioner, CouplingMLP,
                           DAGConditioner, DAGMLP, MADE, MaskedLinear)
from .flow import (FCNormalizingFlow, MNIST_A_prior, NormalLogDensity, NormalizingFlow, NormalizingFlowStep,
                   buildFCNormalizingFlow)
from .normalizers import AffineNormalizer, ELUPlus, IntegrandNet, MonotonicNormalizer, Normalizer
from . import dist
from .graphs import GraphedEvalStep, GraphedTrainStep
from .configs import CONFIGS, build_from_spec
from .image_flows import CIFAR10CNN, CNNormalizingFlow, MNISTCNN, buildCIFAR10NormalizingFlow, buildMNISTNormalizingFlow
from .optim import FusedAdam
from . import experiments
from .experiments import compute_bpp, load_checkpoint, strip_data_parallel_prefix

__all__ = [
    "CIFAR10CNN", "CNNormalizingFlow", "MNISTCNN", "buildCIFAR10NormalizingFlow", "buildMNISTNormalizingFlow",
    "FusedAdam", "experiments", "compute_bpp", "load_checkpoint", "strip_data_parallel_prefix",
    "AutoregressiveConditioner", "Conditioner", "ConditionnalMADE", "CouplingConditioner", "CouplingMLP", "DAGConditioner",
    "DAGMLP", "MADE", "MaskedLinear", "FCNormalizingFlow", "MNIST_A_prior", "NormalLogDensity", "NormalizingFlow",
    "NormalizingFlowStep", "buildFCNormalizingFlow", "AffineNormalizer", "ELUPlus", "IntegrandNet", "MonotonicNormalizer",
    "Normalizer", "ops", "dist", "CONFIGS", "build_from_spec", "GraphedTrainStep", "GraphedEvalStep",
]
