"""Host-side mirror of the reference generator API (/root/reference/model.py:64-132, 206-275).

`Generator(in_channels, out_channels, upscale_factor)` is an nn.Module whose parameters carry exactly the reference
names, shapes, initialisation and RNG consumption order (so `torch.manual_seed(s); Generator(3, 3, 4)` reproduces
the reference weights and reference checkpoints load unchanged), but whose forward is ONE call into the C ABI
(`resr_generator_forward`, include/resr.h) running hand-written sm_100a kernels. The sub-modules exist only to own
parameters; they have no forward of their own — there is no PyTorch fallback path.
"""
import ctypes

import torch
from torch import nn

from . import _lib

__all__ = ["ResidualDenseBlock", "ResidualResidualDenseBlock", "Generator", "plan_tiles", "infer_tiled"]


def _conv(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, (3, 3), (1, 1), (1, 1))


class _ParamsOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise _lib.ResrError(
            f"{type(self).__name__} only owns parameters; run the whole Generator (C ABI resr_generator_forward). "
            "There is no per-block PyTorch path.")


class ResidualDenseBlock(_ParamsOnly):
    """Parameter container of one RDB (reference model.py:64-106): conv1..4 (C+32k -> 32), conv5 (C+128 -> C)."""

    def __init__(self, channels: int, growth_channels: int) -> None:
        super().__init__()
        for k in range(5):
            setattr(self, f"conv{k + 1}",
                    _conv(channels + growth_channels * k, growth_channels if k < 4 else channels))
        # reference model.py:100-106: kaiming_normal * 0.1, zero bias, in module order
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.data *= 0.1
                nn.init.constant_(m.bias, 0)


class ResidualResidualDenseBlock(_ParamsOnly):
    """Parameter container of one RRDB (reference model.py:109-132)."""

    def __init__(self, channels: int, growth_channels: int) -> None:
        super().__init__()
        self.rdb1 = ResidualDenseBlock(channels, growth_channels)
        self.rdb2 = ResidualDenseBlock(channels, growth_channels)
        self.rdb3 = ResidualDenseBlock(channels, growth_channels)


class Generator(nn.Module):
    """RRDBNet x4 (23 RRDB, nf=64, gc=32). Reference: model.py:206-275."""

    def __init__(self, in_channels: int, out_channels: int, upscale_factor: int) -> None:
        super().__init__()
        if (in_channels, out_channels, upscale_factor) != (3, 3, 4):
            # the reference also supports x2/x1 through PixelUnshuffle (model.py:209-217); out of scope
            raise ValueError("resr_b200.Generator implements the x4 RGB configuration Generator(3, 3, 4) only")
        self.conv1 = _conv(in_channels, 64)
        self.trunk = nn.Sequential(*[ResidualResidualDenseBlock(64, 32) for _ in range(23)])
        self.conv2 = _conv(64, 64)
        self.upsampling1 = nn.Sequential(_conv(64, 64), nn.LeakyReLU(0.2, True))
        self.upsampling2 = nn.Sequential(_conv(64, 64), nn.LeakyReLU(0.2, True))
        self.conv3 = nn.Sequential(_conv(64, 64), nn.LeakyReLU(0.2, True))
        self.conv4 = _conv(64, out_channels)
        self._handle = None
        self._packed_version = None
        self._flat = None
        self._workspace = None

    # ------------------------------------------------------------------ native plumbing
    def _native(self):
        if self._handle is None:
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().resr_generator_create(ctypes.byref(h), 3, 3, 4))
            self._handle = h
        return self._handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None:
            try:
                _lib.lib().resr_generator_destroy(h)
            except Exception:
                pass

    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + (next(self.parameters()).device,)

    def flat_parameters(self) -> torch.Tensor:
        """All parameters in state_dict order as one contiguous fp32 device vector (layout of resr_generator_tensor_span)."""
        ps = [p.detach().reshape(-1).float() for p in self.parameters()]
        flat = torch.cat(ps)
        assert flat.numel() == _lib.lib().resr_generator_num_params()
        return flat

    def _ensure_packed(self):
        ver = self._param_version()
        if self._packed_version != ver:
            master = getattr(self, "_flat_master", None)
            first = next(self.parameters())
            if master is not None and first.data_ptr() == master.data_ptr() and first.device == master.device:
                self._flat = master  # parameters are views of one flat vector (optim.FlatAdamEMA): nothing to gather
            else:
                self._flat = self.flat_parameters()
            _lib.check(_lib.lib().resr_generator_load_params(self._native(), _lib.ptr(self._flat), _lib.stream_ptr()))
            self._packed_version = ver

    def _get_workspace(self, n: int, h: int, w: int, device, extra: int = 0) -> torch.Tensor:
        need = _lib.lib().resr_generator_workspace_bytes(n, h, w) + extra
        ws = self._workspace
        if ws is None or ws.numel() < need + 1024 or ws.device != device:
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    @staticmethod
    def _aligned(ws: torch.Tensor):
        base = ws.data_ptr()
        off = (-base) % 1024
        return ctypes.c_void_p(base + off), ws.numel() - off

    # ------------------------------------------------------------------ reference API
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward_impl(x)

    def _forward_impl(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.size(1) != 3:
            raise ValueError(f"expected [N, 3, H, W] input, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise _lib.ResrError("resr_b200.Generator runs on a CUDA (sm_100a) device only; there is no CPU path")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from . import autograd  # training path
            return autograd.generator_apply(self, x)
        return self.infer(x)

    @torch.no_grad()
    def infer(self, x: torch.Tensor) -> torch.Tensor:
        n, _, h, w = x.shape
        xc = x.detach().contiguous(memory_format=torch.contiguous_format).float()
        self._ensure_packed()
        y = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=x.device)
        ws = self._get_workspace(n, h, w, x.device)
        wp, wbytes = self._aligned(ws)
        _lib.check(_lib.lib().resr_generator_forward(self._native(), _lib.ptr(xc), _lib.ptr(y), n, h, w, wp, wbytes,
                                                     _lib.stream_ptr()))
        return y

    @torch.no_grad()
    def infer_host(self, x_host: torch.Tensor, y_host: torch.Tensor = None, device=None) -> torch.Tensor:
        """End-to-end call with HOST tensors (pinned recommended): H2D + forward + D2H inside the C ABI."""
        device = device or next(self.parameters()).device
        n, _, h, w = x_host.shape
        xc = x_host.contiguous().float()
        if y_host is None:
            y_host = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, pin_memory=True)
        self._ensure_packed()
        extra = xc.numel() * 4 * 17 + 4096
        ws = self._get_workspace(n, h, w, device, extra)
        wp, wbytes = self._aligned(ws)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().resr_generator_forward_host(self._native(), _lib.ptr(xc), _lib.ptr(y_host), n, h, w,
                                                              wp, wbytes, _lib.stream_ptr()))
        return y_host


    @torch.no_grad()
    def infer_host_async(self, x_host: torch.Tensor, y_host: torch.Tensor, device=None) -> torch.Tensor:
        """Pipelined serving call: queues H2D + forward + D2H and returns; copies of neighbouring calls overlap this call's
        compute (two staging slots inside the workspace). `x_host` / `y_host` must be pinned fp32 tensors that stay alive
        and untouched until `host_sync()` (or two further calls) — alternate between two `y_host` buffers."""
        device = device or next(self.parameters()).device
        n, _, h, w = x_host.shape
        if not (x_host.is_pinned() and y_host.is_pinned()) or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise _lib.ResrError("infer_host_async needs pinned contiguous fp32 host tensors")
        self._ensure_packed()
        extra = 2 * (x_host.numel() * 4 * 17 + 4096)
        cur = self._workspace
        if cur is not None and (cur.device != device or cur.numel() < _lib.lib().resr_generator_workspace_bytes(n, h, w) + extra + 1024):
            self.host_sync()  # queued copies still use the workspace that is about to be replaced
            torch.cuda.synchronize(device)
        ws = self._get_workspace(n, h, w, device, extra)
        wp, wbytes = self._aligned(ws)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().resr_generator_forward_host_async(self._native(), _lib.ptr(x_host), _lib.ptr(y_host), n, h,
                                                                    w, wp, wbytes, _lib.stream_ptr()))
        return y_host

    def host_sync(self):
        _lib.check(_lib.lib().resr_generator_host_sync(self._native()))


# ----------------------------------------------------------------------------------------------------------------------
# Large-image inference by spatial tiles with a receptive-field halo (BASELINE.json configs[4]; new functionality around
# the unchanged reference API: inference.py:52-53 runs one whole-image forward). Tile indexing is integer arithmetic.


def plan_tiles(h: int, w: int, tile_h: int, tile_w: int, halo: int):
    """Covers an h x w LR image with interior rectangles of at most tile_h x tile_w. Returns a list of
    (y0, y1, x0, x1, wy0, wy1, wx0, wx1): interior [y0,y1) x [x0,x1) and the input window = interior grown by `halo`,
    clamped at the image border (where the convolutions' zero padding applies instead of neighbour data)."""
    tiles = []
    for y0 in range(0, h, tile_h):
        y1 = min(y0 + tile_h, h)
        for x0 in range(0, w, tile_w):
            x1 = min(x0 + tile_w, w)
            tiles.append((y0, y1, x0, x1, max(y0 - halo, 0), min(y1 + halo, h), max(x0 - halo, 0), min(x1 + halo, w)))
    return tiles


@torch.no_grad()
def infer_tiled(gen: "Generator", x: torch.Tensor, tile_h: int = 512, tile_w: int = 1024, halo: int = 16, rank: int = 0,
                world: int = 1, out: torch.Tensor = None):
    """x: [1, 3, H, W] CUDA LR image. Tiles are dealt round-robin to `world` ranks (no collective: every rank computes
    its own tiles from the shared input); this rank's interiors are written into `out` ([1, 3, 4H, 4W], allocated if
    None) and also returned as [(y0, x0, tensor)] for a host-side gather."""
    assert x.dim() == 4 and x.size(0) == 1 and x.size(1) == 3
    _, _, h, w = x.shape
    s = 4
    if out is None:
        out = torch.zeros((1, 3, s * h, s * w), dtype=torch.float32, device=x.device)
    mine = []
    for i, (y0, y1, x0, x1, wy0, wy1, wx0, wx1) in enumerate(plan_tiles(h, w, tile_h, tile_w, halo)):
        if i % world != rank:
            continue
        win = x[:, :, wy0:wy1, wx0:wx1].contiguous()
        sr = gen.infer(win)
        oy, ox = s * (y0 - wy0), s * (x0 - wx0)
        piece = sr[:, :, oy:oy + s * (y1 - y0), ox:ox + s * (x1 - x0)]
        out[:, :, s * y0:s * y1, s * x0:s * x1] = piece
        mine.append((s * y0, s * x0, piece))
    return out, mine
