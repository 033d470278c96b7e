"""``gaussian_renderer.network_gui`` of the drop-in package: the TCP bridge to the SIBR remote viewer.

Not part of the rasterizer hot path, but part of the package's import surface: the reference's
``train_scene.py:17`` does ``from gaussian_renderer import render, network_gui`` and then uses ``network_gui.init``,
``.conn``, ``.try_connect``, ``.receive`` and ``.send`` (``train_scene.py:56-69,228``; reference module:
``gaussian_renderer/network_gui.py:21-85``).  Once the drop-in ``gaussian_renderer`` shadows the reference's package this
module has to exist, with the same module-level state and the viewer's wire protocol: every message is a 4-byte
little-endian length followed by that many bytes; viewer -> trainer messages are UTF-8 JSON camera descriptions, trainer ->
viewer messages are an optional raw image followed by a length-prefixed ASCII tag.
"""
import json
import socket
import traceback

import torch

host = "127.0.0.1"
port = 6009

conn = None        # the connected viewer (None until try_connect() accepted one; the training loop resets it on errors)
addr = None

listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)


def init(wish_host, wish_port):
    """Bind the non-blocking listening socket (train_scene.py:228)."""
    global host, port
    host, port = wish_host, wish_port
    listener.bind((host, port))
    listener.listen()
    listener.settimeout(0)


def try_connect():
    """Accept a waiting viewer, if any; never blocks the training loop."""
    global conn, addr
    try:
        conn, addr = listener.accept()
    except Exception:
        return
    print(f"\nConnected by {addr}")
    conn.settimeout(None)


def _recv_exact(n):
    chunks = []
    while n > 0:
        part = conn.recv(n)
        if not part:
            raise ConnectionError("viewer closed the connection")
        chunks.append(part)
        n -= len(part)
    return b"".join(chunks)


def read():
    """One length-prefixed JSON message from the viewer."""
    size = int.from_bytes(_recv_exact(4), "little")
    return json.loads(_recv_exact(size).decode("utf-8"))


def send(message_bytes, verify):
    """Optional raw payload (the rendered image), then the length-prefixed ASCII tag the viewer checks (the dataset path)."""
    if message_bytes is not None:
        conn.sendall(message_bytes)
    conn.sendall(len(verify).to_bytes(4, "little"))
    conn.sendall(bytes(verify, "ascii"))


def _mini_cam(width, height, fovy, fovx, znear, zfar, world_view_transform, full_proj_transform):
    try:                                   # the reference's own camera type when its `scene` package is importable
        from scene.cameras import MiniCam
        return MiniCam(width, height, fovy, fovx, znear, zfar, world_view_transform, full_proj_transform)
    except ImportError:
        from types import SimpleNamespace  # what gaussian_renderer.render* read from a camera
        return SimpleNamespace(image_width=width, image_height=height, FoVy=fovy, FoVx=fovx, znear=znear, zfar=zfar,
                               world_view_transform=world_view_transform, full_proj_transform=full_proj_transform,
                               camera_center=torch.inverse(world_view_transform)[3][:3])


def receive():
    """-> (camera | None, do_training, convert_SHs_python, compute_cov3D_python, keep_alive, scaling_modifier).
    The viewer's matrices use the opposite y / z handedness: columns 1 and 2 of the view matrix and column 1 of the
    view-projection matrix are negated."""
    msg = read()
    width, height = msg["resolution_x"], msg["resolution_y"]
    if width == 0 or height == 0:
        return None, None, None, None, None, None
    try:
        view = torch.tensor(msg["view_matrix"]).reshape(4, 4).cuda()
        view[:, 1:3] *= -1
        view_proj = torch.tensor(msg["view_projection_matrix"]).reshape(4, 4).cuda()
        view_proj[:, 1] *= -1
        cam = _mini_cam(width, height, msg["fov_y"], msg["fov_x"], msg["z_near"], msg["z_far"], view, view_proj)
        return (cam, bool(msg["train"]), bool(msg["shs_python"]), bool(msg["rot_scale_python"]), bool(msg["keep_alive"]),
                msg["scaling_modifier"])
    except Exception:
        print("")
        traceback.print_exc()
        raise
