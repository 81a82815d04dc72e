"""IoU tracker -- drop-in for the association loop of iouTracke_cal.py (:126-155, flush :174-176; use_iou=True or False)
and its `.npy` output (:177, consumed by iouTracke_display.py:29), computed by fdt_iou_track_metric on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

# module-level defaults of the reference script (iouTracke_cal.py:22-28)
use_iou = True
sigma_iou = 0.4
sigma_dis = 8
sigma_h = 0.6
t_min = 5


def pack_frames(frames):
    """list of per-frame [D_f, 5] arrays (what detect_face returns, iouTracke_cal.py:70-84; float32 rows or
    the float64 dummy [[0,0,0,0,0.4]]) -> (dets[total,5] float64, frame_off[F+1] int64).
    The reference widens to python floats with det0.tolist() (:127), i.e. float64, exactly like this."""
    arrs = [np.asarray(f, dtype=np.float64).reshape(-1, 5) for f in frames]
    off = np.zeros(len(arrs) + 1, np.int64)
    if arrs:
        np.cumsum([a.shape[0] for a in arrs], out=off[1:])
    dets = np.concatenate(arrs, 0) if off[-1] > 0 else np.zeros((1, 5), np.float64)
    return dets, off


def iou_track_raw(dets, frame_off, sigma_iou=sigma_iou, sigma_h=sigma_h, t_min=t_min, use_iou=use_iou, sigma_dis=sigma_dis):
    """dets[total,5] float64 + frame_off[F+1] (numpy or torch, host or device)
    -> (track_off[T+1], track_dets[rows], track_start[T], track_max[T]) as CUDA tensors; tracks are in the
    reference's finishing order, track_dets are global detection rows in append order."""
    dev = _lib.require_cuda()
    d = torch.as_tensor(dets, dtype=torch.float64).to(dev).contiguous()
    off_h = torch.as_tensor(frame_off, dtype=torch.int64).cpu()
    off = off_h.to(dev)
    F = off_h.numel() - 1
    total = int(off_h[-1])
    max_d = int((off_h[1:] - off_h[:-1]).max()) if F > 0 else 0
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    t_off = torch.zeros(total + 2, dtype=torch.int64, device=dev)
    t_dets = torch.zeros(max(total, 1), dtype=torch.int64, device=dev)
    t_start = torch.zeros(total + 1, dtype=torch.int64, device=dev)
    t_max = torch.zeros(total + 1, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L = _lib.lib()
        ws = _lib.workspace(L.fdt_iou_track_workspace_bytes(F, total, max_d), dev, "track")
        _lib.check(L.fdt_iou_track_metric(_lib.ptr(d), _lib.ptr(off), F, total, max_d, 0 if use_iou else 1,
                                          float(sigma_iou if use_iou else sigma_dis), float(sigma_h),
                                          int(t_min), _lib.ptr(n), _lib.ptr(t_off), _lib.ptr(t_dets), _lib.ptr(t_start),
                                          _lib.ptr(t_max), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    T = int(n.item())
    rows = int(t_off[T].item()) if T else 0
    return t_off[:T + 1], t_dets[:rows], t_start[:T], t_max[:T]


def detections_to_frames(detections, width, height, thresh=0.4, shrink=1.0):
    """Detect output of a whole clip, detections [F, C, top_k, 5] on the GPU (one image per frame), -> (dets [total, 5] float64,
    frame_off [F+1] int64), both CUDA tensors, exactly the per-frame arrays detect_face builds (iouTracke_cal.py:55-84: leading rows
    with score >= thresh class by class, boxes * (w, h, w, h) and / shrink in fp32, the dummy row [0, 0, 0, 0, 0.4] for a
    frame without detections) packed back to back: the tracker's input without a round trip through host lists."""
    dev = _lib.require_cuda()
    d = _lib.dev_f32(detections, detections.device if detections.is_cuda else dev)
    F, C_, K, _ = d.shape
    n_rows = torch.empty(max(F, 1), dtype=torch.int32, device=d.device)
    off = torch.empty(F + 1, dtype=torch.int64, device=d.device)
    with torch.cuda.device(d.device):
        L = _lib.lib()
        _lib.check(L.fdt_detections_to_frames_count(_lib.ptr(d), F, C_, K, float(thresh), _lib.ptr(n_rows), _lib.ptr(off), _lib.stream_ptr()))
        total = int(off[-1].item())                      # the one host read: sizes the packed array
        dets = torch.empty((max(total, 1), 5), dtype=torch.float64, device=d.device)
        _lib.check(L.fdt_detections_to_frames_pack(_lib.ptr(d), F, C_, K, float(thresh), float(width), float(height), float(shrink),
                                                   _lib.ptr(off), _lib.ptr(dets), _lib.stream_ptr()))
    return dets[:total] if total else dets[:0], off


def track_detections(detections, width, height, thresh=0.4, shrink=1.0, **kw):
    """detections [F, C, top_k, 5] (CUDA) -> tracks_finished, everything up to the final read-out on the device:
    detections_to_frames + iou_track_raw (iouTracke_cal.py:55-84 feeding :126-155)."""
    dets, off = detections_to_frames(detections, width, height, thresh, shrink)
    t_off, t_dets, t_start, t_max = iou_track_raw(dets, off, **kw)
    t_off, t_dets = t_off.cpu().numpy(), t_dets.cpu().numpy()
    boxes = dets.cpu().numpy()[t_dets, :4] if t_dets.size else np.zeros((0, 4))
    return [{'bboxes': boxes[t_off[t]:t_off[t + 1]].tolist(), 'max_score': float(t_max[t]),
             'start_frame': int(t_start[t])} for t in range(len(t_start))]


def iou_track(frames, sigma_iou=sigma_iou, sigma_h=sigma_h, t_min=t_min, use_iou=use_iou, sigma_dis=sigma_dis):
    """Run the tracker over a whole video.  -> tracks_finished: list of
    {'bboxes': [[x1,y1,x2,y2], ...], 'max_score': float, 'start_frame': int}, the structure
    iouTracke_cal.py:150-154 builds and :177 saves ("track ID" = index in this list)."""
    dets, off = pack_frames(frames)
    t_off, t_dets, t_start, t_max = iou_track_raw(dets, off, sigma_iou, sigma_h, t_min, use_iou, sigma_dis)
    t_off, t_dets = t_off.cpu().numpy(), t_dets.cpu().numpy()
    t_start, t_max = t_start.cpu().numpy(), t_max.cpu().numpy()
    boxes = dets[t_dets, :4]
    return [{'bboxes': boxes[t_off[t]:t_off[t + 1]].tolist(), 'max_score': float(t_max[t]),
             'start_frame': int(t_start[t])} for t in range(len(t_start))]


def save_tracks(path, tracks_finished):
    """np.save(video_file + ".npy", np.array(tracks_finished))   [iouTracke_cal.py:177] -- a pickled 1-D object
    array of dicts; `path` gets the .npy suffix from numpy exactly as in the reference."""
    arr = np.empty(len(tracks_finished), dtype=object)
    for i, t in enumerate(tracks_finished):
        arr[i] = t
    np.save(path, arr, allow_pickle=True)


def load_tracks(path):
    """tracks = np.load(video_file + '.npy').tolist()   [iouTracke_display.py:29]; modern numpy needs allow_pickle."""
    return np.load(path, allow_pickle=True).tolist()
