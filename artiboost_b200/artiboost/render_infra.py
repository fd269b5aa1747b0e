"""Render provider (anakin/artiboost/render_infra.py:62-152): same constructor, queues and message protocol.

Request  = {"id": worker id, "objname": str, "pose": f32[4,4], "hand_verts": f32[778,3]} on `get_message_queue()`
Reply    = uint8[H,W,3] BGR ndarray on `get_image_queue_list()[id]`            (rendered_dataset.py:118-123)

The reference forks one pyrender/EGL process per render GPU, each serving ONE image per queue round trip.  Here one
server thread per render GPU drains up to `max_batch` pending requests and rasterises them in a single
ab_render_batch call.  On-device callers skip the queues and use `provider.renderers[i].render_batch` directly.
"""
import atexit
import queue as pyqueue
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.multiprocessing as mp

from . import DUMMY
from .renderer import Renderer


def render_worker(exit_event, renderer: Renderer, incoming_queue, output_queue_list, max_batch: int):
    """Server loop: same polling/exit protocol as render_infra.py:46-59, but batched."""
    dev = renderer.device
    while not exit_event.is_set():
        try:
            msgs = [incoming_queue.get(block=True, timeout=0.2)]
        except pyqueue.Empty:
            continue
        while len(msgs) < max_batch:
            try:
                msgs.append(incoming_queue.get_nowait())
            except pyqueue.Empty:
                break
        ids = [-1 if m["objname"] == DUMMY else renderer.obj_index[m["objname"]] for m in msgs]
        pose = np.stack([np.eye(4, dtype=np.float32) if i < 0 else np.asarray(m["pose"], np.float32).reshape(4, 4)
                         for i, m in zip(ids, msgs)])
        hv = np.stack([np.asarray(m["hand_verts"], np.float32) for m in msgs])
        out = renderer.render_batch(torch.tensor(ids, dtype=torch.int32, device=dev), torch.from_numpy(pose).to(dev),
                                    torch.from_numpy(hv).to(dev), want=("rgba",))
        bgr = out["rgba"][..., :3].flip(-1).cpu().numpy()
        for m, img in zip(msgs, bgr):
            output_queue_list[m["id"]].put(img)
        del msgs


class RendererProvider:

    def __init__(self, num_workers: int, gpu_render_id: Sequence[int], render_size: List[int], cam_intr: np.ndarray,
                 cam_extr: np.ndarray, obj_meshes: Dict[str, object], hand_meshes: List[object],
                 bgs: Optional[List[object]] = None, lights: Optional[List[object]] = None, cfg_renderer=None,
                 cfg_datapreset=None, arg_extra=None, random_seed=1, max_batch: int = 64):
        self.gpu_render_id = list(gpu_render_id)
        self.gpu_render_used = len(self.gpu_render_id)
        self.exit_event = mp.Event()
        self.message_queue = mp.Queue()
        self.image_queue_list = [mp.Queue() for _ in range(num_workers)]
        self.renderers: List[Renderer] = []
        self.server_proc_list: List[threading.Thread] = []
        for proc_id, gpu_id in enumerate(self.gpu_render_id):
            r = Renderer(width=render_size[0], height=render_size[1], gpu_id=gpu_id)
            r.rng = np.random.RandomState(random_seed + proc_id)  # render_infra.py:30-31
            r.setup(cam_intr=cam_intr, cam_extr=cam_extr, obj_meshes=obj_meshes, hand_meshes=hand_meshes,
                    backgrounds=bgs, lights=lights)
            self.renderers.append(r)
            self.server_proc_list.append(threading.Thread(
                target=render_worker, daemon=True,
                kwargs=dict(exit_event=self.exit_event, renderer=r, incoming_queue=self.message_queue,
                            output_queue_list=self.image_queue_list, max_batch=max_batch)))
        self.running = False

        def gracefully_exit_fn():
            if self.running:
                self.exit_event.set()
                for t in self.server_proc_list:
                    t.join()
                self.running = False

        self.gracefully_exit = gracefully_exit_fn
        atexit.register(self.gracefully_exit)

    def __del__(self):
        self.gracefully_exit()

    def begin(self):
        if not self.running:
            self.running = True
            for t in self.server_proc_list:
                t.start()

    def end(self):
        if self.running:
            self.gracefully_exit()

    def get_process_list(self):
        return self.exit_event

    def get_message_queue(self):
        return self.message_queue

    def get_image_queue_list(self):
        return self.image_queue_list
