"""Depth-image VAE encoder used by NavigationTask (utils/vae/vae_image_encoder.py:19-56 wrapping the
encoder of utils/vae/VAE.py:67-155).  Stays in torch: cuDNN convolutions on a
[N,1,270,480] image are not part of the hand-written hot paths.

The layer table below reproduces the reference encoder's topology and parameter names
(`encoder.conv0.weight`, ...), so the reference checkpoint loads with `load_state_dict`; the decoder
half of the checkpoint is ignored."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

# name, in, out, kernel, stride, padding
_CONVS = (
    ("conv0", 1, 32, 5, 2, 2), ("conv0_1", 32, 32, 3, 2, 2),
    ("conv1_0", 32, 32, 5, 2, 1), ("conv1_1", 32, 64, 3, 1, 1),
    ("conv2_0", 64, 64, 5, 2, 2), ("conv2_1", 64, 128, 3, 2, 1),
    ("conv3_0", 128, 128, 5, 2, 0),
    ("conv0_jump_2", 32, 64, 4, 2, 1), ("conv1_jump_3", 64, 128, 5, 4, (2, 1)),
)


class _Encoder(nn.Module):
    def __init__(self, latent_dim):
        super().__init__()
        for name, cin, cout, k, s, p in _CONVS:
            setattr(self, name, nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p))
        self.dense0 = nn.Linear(3 * 6 * 128, 512)
        self.dense1 = nn.Linear(512, 2 * latent_dim)

    def forward(self, img):
        a = F.elu(self.conv0_1(self.conv0(img)))
        b = F.elu(self.conv1_1(self.conv1_0(a)) + self.conv0_jump_2(a))   # first residual jump
        c = F.elu(self.conv2_1(self.conv2_0(b)) + self.conv1_jump_3(b))   # second residual jump
        x = self.conv3_0(c).flatten(1)
        return self.dense1(F.elu(self.dense0(x)))


class VAEImageEncoder(nn.Module):
    def __init__(self, config, device="cuda:0"):
        super().__init__()
        self.config = config
        self.latent_dim = int(config.latent_dims)
        self.encoder = _Encoder(self.latent_dim).to(device)
        path = os.path.join(config.model_folder, config.model_file) if config.model_file else ""
        self.weights_loaded = False
        if path and os.path.isfile(path):
            sd = torch.load(path, map_location=device)
            clean = {}
            for k, v in sd.items():  # vae_image_encoder.py:8-16
                k = k.replace("module.", "").replace("dronet.", "encoder.")
                if k.startswith("encoder."):
                    clean[k[len("encoder."):]] = v
            self.encoder.load_state_dict(clean)
            self.weights_loaded = True
        self.encoder.eval()
        # CUDA: NHWC weights / activations and cuDNN's autotuner (cudnn.benchmark) -- the same convolutions, the fastest algorithm cuDNN
        # has for each shape (the encoder is 83 % of a navigation_task env step at 1024 envs: bench.py navigation_task_e2e.breakdown)
        self._cuda = torch.device(device).type == "cuda"
        if self._cuda:
            self.encoder.to(memory_format=torch.channels_last)
            torch.backends.cudnn.benchmark = True

    @torch.no_grad()
    def encode(self, image_tensors):
        x = image_tensors.squeeze(0).unsqueeze(1)
        if tuple(self.config.image_res) != tuple(x.shape[-2:]):
            x = F.interpolate(x, tuple(self.config.image_res), mode=self.config.interpolation_mode)
        if self._cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        z = self.encoder(x)
        means, logvars = z[:, : self.latent_dim], z[:, self.latent_dim:]
        if not self.config.return_sampled_latent:
            return means
        return means + torch.randn_like(logvars) * torch.exp(0.5 * logvars)  # VAE.encode, VAE.py:228-244

    forward = encode
