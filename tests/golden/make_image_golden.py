"""Golden vectors of the reference's image-flow path: DAGConditioner(hidden=MNISTCNN) steps stacked in a
two-scale CNNormalizingFlow (14x14 -> 7x7, the second and third scales of buildMNISTNormalizingFlow,
models/NormalizingFlowFactories.py:49-80), run by the UNMODIFIED reference in the build container:

    python tests/golden/make_image_golden.py    ->  tests/golden/image_cnn_flow.npz

Deterministic gates (stoch_gate = False) so that no noise has to be replayed; inputs, state_dict, z, log-det, loss and every gradient."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")
import networkx as nx  # noqa: E402

if not hasattr(nx, "from_numpy_matrix"):
    nx.from_numpy_matrix = nx.from_numpy_array

from models.Conditionners import DAGConditioner  # noqa: E402
from models.MLP import MNISTCNN  # noqa: E402
from models.Normalizers import AffineNormalizer, MonotonicNormalizer  # noqa: E402
from models.NormalizingFlow import CNNormalizingFlow, FCNormalizingFlow, NormalizingFlowStep  # noqa: E402
from models.NormalizingFlowFactories import MNIST_A_prior, NormalLogDensity  # noqa: E402


def build(classes, hot=False):
    DAG, CNN, Aff, Mono, Step, FC, CNF, prior, Dens = classes
    torch.manual_seed(3)
    outer = []
    for img, fc, ntype in (([1, 14, 14], [400, 64], "affine"), ([1, 7, 7], [16, 16], "monotonic")):
        d = img[1] * img[2]
        emb = 2 if ntype == "affine" else 6
        cond = DAG(d, CNN(fc_l=fc, size_img=img, out_d=emb), emb, l1=.3, nb_epoch_update=10, hot_encoding=False, A_prior=prior(img[1], 2))
        norm = Aff() if ntype == "affine" else Mono(integrand_net=[20, 20], cond_size=emb, nb_steps=12, solver="CC")
        flow = FC([Step(cond, norm)], None)
        flow.img_sizes = img
        outer.append(flow)
    return CNF(outer, Dens(), [[1, 2, 2], [1, 1, 1]])


def main():
    model = build((DAGConditioner, MNISTCNN, AffineNormalizer, MonotonicNormalizer, NormalizingFlowStep, FCNormalizingFlow, CNNormalizingFlow,
                   MNIST_A_prior, NormalLogDensity))
    with torch.no_grad():                       # a non-binary adjacency inside the prior's support, so that dA is informative
        for c in model.getConditioners():
            c.A.mul_(.5 + torch.rand_like(c.A))
    for c in model.getConditioners():
        c.stoch_gate = False
    x = torch.randn(3, 196, generator=torch.Generator().manual_seed(9))
    z, jac = model(x)
    loss = model.loss(z, jac)
    loss.backward()
    out = {"x": x.numpy(), "z": z.detach().numpy(), "logdet": jac.detach().numpy(), "loss": np.float32(float(loss.detach()))}
    for k, v in model.state_dict().items():
        out["sd." + k] = v.detach().numpy()
    for k, p in model.named_parameters():
        if p.grad is not None:
            out["grad." + k] = p.grad.numpy()
    path = os.path.join(HERE, "image_cnn_flow.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len([k for k in out if k.startswith('grad.')]), "gradients; loss", float(loss))


if __name__ == "__main__":
    main()
