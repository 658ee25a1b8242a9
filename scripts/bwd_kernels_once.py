"""One ACE-sized backward of a 1x1 convolution (512 <- 256 channels) and of a fused spectral convolution (256 channels,
180x360, dhconv), batch 4, bf16: the launches ncu captures for the weight-gradient kernels."""
import sys

sys.path.insert(0, ".")
import torch

import spherical_dyffusion_b200 as sb
from spherical_dyffusion_b200 import _lib

dev = torch.device("cuda:0")
B, C, H, W = 4, 256, 180, 360
x = torch.randn(B, C, H, W, device=dev, requires_grad=True)
w = torch.randn(512, C, 1, 1, device=dev, requires_grad=True)
b = torch.zeros(512, device=dev, requires_grad=True)
y = torch.ops.sfno_b200.conv1x1_ex(x, w, b, None, 0, 0.0, 0, 0, _lib.SFNO_PREC["bf16"])
y.sum().backward()
sht = sb.RealSHT(H, W, lmax=180, mmax=181, grid="legendre-gauss", precision="bf16")
isht = sb.InverseRealSHT(H, W, lmax=180, mmax=181, grid="legendre-gauss", precision="bf16")
conv = sb.SpectralConvS2(sht, isht, C, C, operator_type="dhconv", bias=True).to(dev)
x2 = torch.randn(B, C, H, W, device=dev, requires_grad=True)
out, _ = conv(x2)
out.sum().backward()
torch.cuda.synchronize()
print("ok", float(conv.weight.grad.abs().sum()))
