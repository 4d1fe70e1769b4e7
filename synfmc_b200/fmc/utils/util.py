"""Host-side helpers the trainers import from `fmc.utils.util` (train_cam_ctrl.py:32, train_cam_obj_ctrl.py:32):
`setup_logger`, `format_time`, `save_videos_grid`, plus `instantiate_from_config` (fmc/utils/util.py:16-33).  None of
this is on the denoising path; it exists so that the mirror package satisfies the trainers' imports."""
import atexit
import importlib
import logging
import os
import sys

import numpy as np
import torch

_UNITS = ((86400, "days"), (3600, "hours"), (60, "minutes"))


def format_time(elapsed_time):
    """Seconds -> 'D days H hours M minutes S.SS seconds', zero-valued parts dropped (fmc/utils/util.py:127-149)."""
    rest = elapsed_time
    parts = []
    for span, label in _UNITS:
        count, rest = divmod(rest, span)
        if count > 0:
            parts.append(f"{int(count)} {label}")
    if rest > 0:
        parts.append(f"{rest:.2f} seconds")
    return " ".join(parts)


def get_obj_from_str(string, reload=False):
    module_name, attr = string.rsplit(".", 1)
    module = importlib.import_module(module_name)
    if reload:
        module = importlib.reload(module)
    return getattr(module, attr)


def instantiate_from_config(config, **additional_kwargs):
    """{'target': 'pkg.mod.Class', 'kwargs': {...}} -> Class(**kwargs) (fmc/utils/util.py:16-25)."""
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    additional_kwargs.update(config.get("kwargs", dict()))
    return get_obj_from_str(config["target"])(**additional_kwargs)


_STREAMS = {}
_LOGGERS = {}


def _log_stream(filename):
    """one append-mode stream per file name, shared by every logger that writes there"""
    if filename not in _STREAMS:
        _STREAMS[filename] = open(filename, "a", buffering=1024 if "://" in filename else -1)
        atexit.register(_STREAMS[filename].close)
    return _STREAMS[filename]


class _LevelPrefixFormatter(logging.Formatter):
    """console format: abbreviated logger name, WARNING / ERROR prefixed in red (ANSI), everything else plain"""

    def __init__(self, fmt, datefmt, root_name, abbrev_name):
        super().__init__(fmt, datefmt=datefmt)
        self._root = root_name + "."
        self._abbrev = (abbrev_name + ".") if abbrev_name else ""

    def formatMessage(self, record):
        record.name = record.name.replace(self._root, self._abbrev)
        line = super().formatMessage(record)
        if record.levelno == logging.WARNING:
            return "\033[5m\033[31mWARNING\033[0m " + line
        if record.levelno >= logging.ERROR:
            return "\033[4m\033[5m\033[31mERROR\033[0m " + line
        return line


def setup_logger(output, distributed_rank, color=True, name="AnimateDiff", abbrev_name=None):
    """Logger `name` at DEBUG, not propagating: stdout handler on rank 0 only; file handler on every rank writing
    `output` itself if it ends in .txt / .log, else `output/log.txt`, with `.rank<N>` appended for N > 0.  Calls with
    the same arguments return the same logger without adding handlers again (fmc/utils/util.py:82-124)."""
    key = (output, distributed_rank, color, name, abbrev_name)
    if key in _LOGGERS:
        return _LOGGERS[key]
    logger = logging.getLogger(name)
    logger.setLevel(logging.DEBUG)
    logger.propagate = False
    plain = logging.Formatter("[%(asctime)s] %(name)s:%(lineno)d %(levelname)s: %(message)s", datefmt="%m/%d %H:%M:%S")
    if distributed_rank == 0:
        console = logging.StreamHandler(stream=sys.stdout)
        console.setLevel(logging.DEBUG)
        if color:
            console.setFormatter(_LevelPrefixFormatter("\033[32m[%(asctime)s %(name)s:%(lineno)d]: \033[0m%(message)s",
                                                       "%m/%d %H:%M:%S", name, str(abbrev_name or "AD")))
        else:
            console.setFormatter(plain)
        logger.addHandler(console)
    if output is not None:
        filename = output if output.endswith((".txt", ".log")) else os.path.join(output, "log.txt")
        if distributed_rank > 0:
            filename = f"{filename}.rank{distributed_rank}"
        os.makedirs(os.path.dirname(filename), exist_ok=True)
        to_file = logging.StreamHandler(_log_stream(filename))
        to_file.setLevel(logging.DEBUG)
        to_file.setFormatter(plain)
        logger.addHandler(to_file)
    _LOGGERS[key] = logger
    return logger


def _frame_grid(frame, n_rows, padding=2):
    """[b, c, h, w] -> [c, H, W] tiles, `n_rows` images per row, 2-pixel zero border: the layout of
    torchvision.utils.make_grid(frame, nrow=n_rows) with its defaults, which fmc/utils/util.py:40 calls (a single image
    is returned as it is, like make_grid does)."""
    b, c, h, w = frame.shape
    if c == 1:
        frame = frame.expand(b, 3, h, w)
        c = 3
    if b == 1:
        return frame[0]
    per_row = min(n_rows, b)
    n_lines = -(-b // per_row)
    grid = frame.new_zeros(c, n_lines * (h + padding) + padding, per_row * (w + padding) + padding)
    for k in range(b):
        y, x = divmod(k, per_row)
        grid[:, y * (h + padding) + padding: y * (h + padding) + padding + h,
             x * (w + padding) + padding: x * (w + padding) + padding + w] = frame[k]
    return grid


def video_grid_frames(videos, rescale=False, n_rows=6):
    """[b, c, t, h, w] in [0, 1] (or [-1, 1] with rescale) -> list of t uint8 [H, W, 3] frames (truncating cast, as the
    reference's `(x * 255).numpy().astype(np.uint8)`, fmc/utils/util.py:36-46)."""
    frames = []
    for i in range(videos.shape[2]):
        x = _frame_grid(videos[:, :, i].detach().cpu().float(), n_rows).permute(1, 2, 0)
        if rescale:
            x = (x + 1.0) / 2.0
        frames.append((x * 255).numpy().astype(np.uint8))
    return frames


def save_videos_grid(videos: torch.Tensor, path: str, rescale=False, n_rows=6, fps=8):
    """Write the frame grids as an animation at `path` (fmc/utils/util.py:36-50; the reference goes through
    imageio.mimsave).  imageio is used when it is installed; otherwise .gif files are written with Pillow, and any
    other container needs imageio."""
    frames = video_grid_frames(videos, rescale=rescale, n_rows=n_rows)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    try:
        import imageio
    except ImportError:
        imageio = None
    if imageio is not None:
        imageio.mimsave(path, frames, fps=fps)
        return
    if not path.lower().endswith(".gif"):
        raise RuntimeError(f"save_videos_grid: writing {os.path.splitext(path)[1]} needs imageio (not installed); .gif works without it")
    from PIL import Image
    images = [Image.fromarray(f) for f in frames]
    images[0].save(path, save_all=True, append_images=images[1:], duration=int(round(1000 / fps)), loop=0)
