"""The 3 x 21 icosahedral positional-encoding basis of Mip-NeRF 360 (`generate_basis("icosahedron", 2)`,
models/mipnerf360/helper.py:457-531, registered as the buffer `pos_basis_t` of every MipNeRF360MLP, model.py:66-68).
It is init-time numpy in the reference; the 63 numbers are embedded here (fp32) and checked against the reference's
generator by oracle/make_golden.py."""
import torch

_A, _B, _C, _D = 0.8506507873535156, 0.8090170025825500, 0.5257310867309570, 0.3090170025825500

POS_BASIS_T = torch.tensor([
    [_A, _B, _C, 1.0, _B, _A, _D, 0.0, 0.5, 0.0, -_C, -_D, 0.0, -_D, _D, 0.5, 0.5, 0.0, -0.5, -_B, -_B],
    [0.0, 0.5, _A, 0.0, 0.5, 0.0, _B, _C, _D, 1.0, _A, _B, _C, _B, _B, _D, -_D, 0.0, _D, 0.5, 0.5],
    [_C, _D, 0.0, 0.0, -_D, -_C, -0.5, -_A, -_B, 0.0, 0.0, -0.5, _A, 0.5, 0.5, _B, _B, 1.0, _B, _D, -_D],
], dtype=torch.float32)
