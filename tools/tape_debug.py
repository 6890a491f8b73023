"""B200_TAPE_DEBUG=1 python tools/tape_debug.py: where every bias gradient of a small ResU-Net pass comes from (Tape.grad_sums / memo / a pass
over dy) and which normalisation backward could state the channel sums of its input gradient."""
import contextlib, io, torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biapy_b200.models.resunet import ResUNet
kw = dict(image_shape=(32, 32, 128, 2), activation="silu", feature_maps=[16, 32, 64], drop_values=[0, 0, 0],
          normalization="gn", k_size=3, yx_down=[2, 2], z_down=[2, 2], isotropy=[True] * 3, larger_io=False,
          conv_layers=[2] * 3, output_channels=[1])
with contextlib.redirect_stdout(io.StringIO()):
    model = ResUNet(**kw).cuda()
model.set_engine(dtype=torch.float16); model.train()
x = torch.randn(1, 2, 32, 32, 128, device="cuda")
y = model(x); y.sum().backward(); torch.cuda.synchronize()
