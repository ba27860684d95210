"""Accuracy of quant_math.cuh::gelu_fast (restated in numpy fp32) against fp64 erf, next to torch's fp32 F.gelu."""
import math

import numpy as np
import torch
from scipy.special import erf

f = np.float32
g = np.linspace(-8, 8, 2000001).astype(f)
ref = 0.5 * g.astype(np.float64) * (1 + erf(g.astype(np.float64) / math.sqrt(2)))
p = f(0.3275911 * 0.7071067811865476)
a = [f(0.5 * x) for x in (0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429)]
ag = np.abs(g)
t = (f(1) / (ag * p + f(1))).astype(f)
poly = ((((a[4] * t + a[3]) * t + a[2]) * t + a[1]) * t + a[0])
e = np.exp2((g * f(-0.5 * 1.4426950408889634)) * g).astype(f)
h = (ag * ((poly * t) * e)).astype(f)
out = np.where(g >= 0, g - h, -h).astype(f)
print("gelu_fast  max |err| %.3e" % np.abs(out.astype(np.float64) - ref).max())
print("torch fp32 max |err| %.3e" % np.abs(torch.nn.functional.gelu(torch.from_numpy(g)).numpy().astype(np.float64) - ref).max())
