"""Restatement of the reference's cost-regularisation 3-D U-Net, ``CostRegNet_3DGS``.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (same rules as mvsdet_oracle.py).  The net sits BETWEEN the
stages of the hot path (mvsdet.py:470) and is not part of it; it is restated here only so that the
GPU tests, which cannot read /root/reference, can attach the real architecture to the drop-in
(the variance hand-off to cuDNN).  Follows
projects/NeRF-Det/nerfdet/mvs_models/mvsnet.py:73-113 and ConvBnReLU3D (module.py:26-33) with the
SAME module and parameter names, so a reference ``state_dict`` loads unchanged;
tests/test_oracle_vs_reference.py checks key / shape identity and output equality against the
reference class on the build host.
"""
import torch.nn as nn
import torch.nn.functional as F


class ConvBnReLU3D(nn.Module):                          # module.py:26-33
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)


def _up(cin, cout):                                     # mvsnet.py:91-99
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
        nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class CostRegNet3DGS(nn.Module):
    """[V,256,D,H,W] variance volume -> [V,2,D,H,W] (cost, offset); D, H, W divisible by 4."""

    def __init__(self, in_channels: int = 256):
        super().__init__()
        self.conv0 = ConvBnReLU3D(in_channels, 64)      # mvsnet.py:76
        self.conv1 = ConvBnReLU3D(64, 128, stride=2)
        self.conv2 = ConvBnReLU3D(128, 128)
        self.conv3 = ConvBnReLU3D(128, 256, stride=2)
        self.conv4 = ConvBnReLU3D(256, 256)
        self.conv9 = _up(256, 128)
        self.conv11 = _up(128, 64)
        self.prob = nn.Conv3d(64, 2, 3, stride=1, padding=1)

    def forward(self, x):                               # mvsnet.py:103-113
        conv0 = self.conv0(x)
        conv2 = self.conv2(self.conv1(conv0))
        x = self.conv4(self.conv3(conv2))
        x = conv2 + self.conv9(x)
        x = conv0 + self.conv11(x)
        return self.prob(x)
