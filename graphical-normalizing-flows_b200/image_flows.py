"""Image flows of the reference: DAG conditioners whose embedding network is a small CNN over the masked
image (models/MLP.py:24-72), stacked in a multi-scale `CNNormalizingFlow` (models/NormalizingFlow.py:172-226) by the MNIST /
CIFAR-10 factories (models/NormalizingFlowFactories.py:49-135).

What is B200-native here is the DAG part: the gated masked copies of the input ([B, d, d], stochastic Gumbel gate included) come
from the fused DAG layer-1 kernels and the normalizers / log-det / base density run on the hand-written kernels.  The CNN body
itself (two convolutions on B*d tiny images) is the user's nn.Module and runs as given (cuDNN), exactly like any other
`hidden=<nn.Module>`.  Class and parameter names follow the reference so that its checkpoints load."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .conditioners import DAGConditioner
from .flow import FCNormalizingFlow, MNIST_A_prior, NormalLogDensity, NormalizingFlowStep
from .normalizers import AffineNormalizer, MonotonicNormalizer


class MNISTCNN(nn.Module):
    """conv3x3(16) - relu - conv3x3(16) - maxpool2 - fc - relu - fc  (models/MLP.py:24-48; its dropouts are disabled there too)."""

    def __init__(self, out_d=10, fc_l=(2304, 128), size_img=(1, 28, 28)):
        super().__init__()
        self.conv1 = nn.Conv2d(size_img[0], 16, 3, 1)
        self.conv2 = nn.Conv2d(16, 16, 3, 1)
        self.dropout1 = nn.Dropout2d(0.25)
        self.dropout2 = nn.Dropout2d(0.5)
        self.fc1 = nn.Linear(fc_l[0], fc_l[1])
        self.fc2 = nn.Linear(fc_l[1], out_d)
        self.out_d, self.size_img = out_d, list(size_img)

    def forward(self, x, context=None):
        n = x.shape[0]
        y = F.relu(self.conv1(x.view(-1, *self.size_img)))
        y = torch.flatten(F.max_pool2d(self.conv2(y), 2), 1)
        return self.fc2(F.relu(self.fc1(y))).view(n, -1)


class CIFAR10CNN(nn.Module):
    """conv(6) - relu - pool - conv(16) - relu - pool - fc - relu - fc - relu - fc  (models/MLP.py:51-72)."""

    def __init__(self, out_d=10, fc_l=(400, 128, 84), size_img=(3, 32, 32), k_size=5):
        super().__init__()
        self.conv1 = nn.Conv2d(size_img[0], 6, k_size)
        self.pool = nn.MaxPool2d(2, 2)
        self.conv2 = nn.Conv2d(6, 16, k_size)
        self.fc1 = nn.Linear(fc_l[0], fc_l[1])
        self.fc2 = nn.Linear(fc_l[1], fc_l[2])
        self.fc3 = nn.Linear(fc_l[2], out_d)
        self.out_d, self.size_img = out_d, list(size_img)

    def forward(self, x, context=None):
        n = x.shape[0]
        y = self.pool(F.relu(self.conv1(x.view(-1, *self.size_img))))
        y = self.pool(F.relu(self.conv2(y))).view(n, -1)
        return self.fc3(F.relu(self.fc2(F.relu(self.fc1(y))))).view(n, -1)


def _squeeze_blocks(z, img, drop):
    """[B, C*H*W] -> [B, c, h, w, dc*dh*dw]: the image cut into dc x dh x dw blocks, block content last (the reference's triple unfold)."""
    (C, H, W), (dc, dh, dw) = img, drop
    c, h, w = C // dc, H // dh, W // dw
    t = z.view(-1, c, dc, h, dh, w, dw).permute(0, 1, 3, 5, 2, 4, 6)
    return t.reshape(z.shape[0], c, h, w, dc * dh * dw)


class CNNormalizingFlow(FCNormalizingFlow):
    """Multi-scale flow (NormalizingFlow.py:172-226): after every inner flow the image is cut into blocks; the first element of
    each block goes on to the next (smaller) scale, the others are emitted as latents."""

    def __init__(self, steps, z_log_density, dropping_factors):
        super().__init__(steps, z_log_density)
        self.dropping_factors = dropping_factors

    def forward(self, x, context=None):
        B = x.shape[0]
        jac_tot, latents = 0., []
        for step, drop in zip(self.steps, self.dropping_factors):
            z, jac = step(x, context)
            blocks = _squeeze_blocks(z, step.img_sizes, drop)
            latents.append(blocks[..., 1:].reshape(B, -1))
            x = blocks[..., 0].reshape(B, -1)
            jac_tot = jac_tot + jac
        latents.append(x)
        return torch.cat(latents, 1), jac_tot

    def invert(self, z, context=None):
        B = z.shape[0]
        parts, i = [], 0
        for step, (dc, dh, dw) in zip(self.steps, self.dropping_factors):
            C, H, W = step.img_sizes
            kept = (C // dc) * (H // dh) * (W // dw)
            n = C * H * W - kept if C * H * W != kept else kept
            parts.append(z[:, i:i + n])
            i += n
        x = None
        for step, (dc, dh, dw), part in zip(reversed(list(self.steps)), reversed(list(self.dropping_factors)), reversed(parts)):
            C, H, W = step.img_sizes
            c, h, w = C // dc, H // dh, W // dw
            if c * h * w != C * H * W:
                blocks = torch.cat((x.view(B, c, h, w, 1), part.reshape(B, c, h, w, -1)), 4).view(B, c, h, w, dc, dh, dw)
                part = blocks.permute(0, 1, 4, 2, 5, 3, 6).reshape(B, C * H * W)
            x = step.invert(part.reshape(B, -1), context)
        return x


def _dag_cnn_step(in_size, cnn, emb, normalizer_type, normalizer_args, l1, nb_epoch_update, hot_encoding, A_prior, mono_cond_size):
    cond = DAGConditioner(in_size, cnn, emb, l1=l1, nb_epoch_update=nb_epoch_update, hot_encoding=hot_encoding, A_prior=A_prior)
    if normalizer_type is MonotonicNormalizer and mono_cond_size is not None:
        norm = normalizer_type(**normalizer_args, cond_size=mono_cond_size)
    else:
        norm = normalizer_type(**normalizer_args)
    return NormalizingFlowStep(cond, norm)


def buildMNISTNormalizingFlow(nb_inner_steps, normalizer_type, normalizer_args, l1=0., nb_epoch_update=10, hot_encoding=False,
                              prior_kernel=None):
    """NormalizingFlowFactories.py:49-97.  Three scales (28, 14, 7) or a single 28 x 28 flow; anything else returns None."""
    emb = 2 if normalizer_type is AffineNormalizer else 30

    def steps_for(img, fc, count):
        d = img[0] * img[1] * img[2]
        out = []
        for _ in range(count):
            prior = MNIST_A_prior(img[1], prior_kernel) if prior_kernel is not None else None
            out.append(_dag_cnn_step(d, MNISTCNN(fc_l=fc, size_img=img, out_d=emb), emb, normalizer_type, normalizer_args, l1,
                                     nb_epoch_update, hot_encoding, prior, (30 + d) if hot_encoding else 30))
        return out

    if len(nb_inner_steps) == 3:
        scales = zip([[1, 28, 28], [1, 14, 14], [1, 7, 7]], [[2304, 128], [400, 64], [16, 16]], nb_inner_steps)
        outer = []
        for img, fc, count in scales:
            flow = FCNormalizingFlow(steps_for(img, fc, count), None)
            flow.img_sizes = img
            outer.append(flow)
        return CNNormalizingFlow(outer, NormalLogDensity(), [[1, 2, 2], [1, 2, 2], [1, 1, 1]])
    if len(nb_inner_steps) == 1:
        return FCNormalizingFlow(steps_for([1, 28, 28], [2304, 128], nb_inner_steps[0]), NormalLogDensity())
    return None


def buildCIFAR10NormalizingFlow(nb_inner_steps, normalizer_type, normalizer_args, l1=0., nb_epoch_update=5):
    """NormalizingFlowFactories.py:100-135.  Four scales (3x32x32, 32, 16, 8) or a single 3x32x32 flow; anything else returns None."""
    emb = 2 if normalizer_type is AffineNormalizer else 30

    def steps_for(img, fc, k, count):
        d = img[0] * img[1] * img[2]
        return [_dag_cnn_step(d, CIFAR10CNN(out_d=emb, fc_l=fc, size_img=img, k_size=k), emb, normalizer_type, normalizer_args, l1,
                              nb_epoch_update, False, None, None) for _ in range(count)]

    if len(nb_inner_steps) == 4:
        imgs = [[3, 32, 32], [1, 32, 32], [1, 16, 16], [1, 8, 8]]
        fcs = [[400, 128, 84], [576, 128, 32], [64, 32, 32], [16, 32, 32]]
        outer = []
        for img, fc, k, count in zip(imgs, fcs, [5, 3, 3, 2], nb_inner_steps):
            flow = FCNormalizingFlow(steps_for(img, fc, k, count), None)
            flow.img_sizes = img
            outer.append(flow)
        return CNNormalizingFlow(outer, NormalLogDensity(), [[3, 1, 1], [1, 2, 2], [1, 2, 2]])
    if len(nb_inner_steps) == 1:
        return FCNormalizingFlow(steps_for([3, 32, 32], [400, 128, 84], 5, nb_inner_steps[0]), NormalLogDensity())
    return None
