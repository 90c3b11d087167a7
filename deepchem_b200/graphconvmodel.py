"""GraphConvModel: the drop-in model class for the B200 path.

Mirrors deepchem/models/torch_models/graphconvmodel.py (``TrimGraphOutput`` :21,
``_GraphConvTorchModel`` :36, ``GraphConvModel`` :252) and the parts of ``TorchModel``
(deepchem/models/torch_models/torch_model.py) the path goes through: ``fit`` / ``fit_generator``
(:289-496), ``fit_on_batch``, ``predict`` / ``predict_on_batch`` / ``predict_embedding``
(:547-761), ``evaluate``, ``save_checkpoint`` / ``restore`` (:996-1090).

Differences from the reference torch port, each one deliberate:
  * gradients flow through GraphConv / GraphPool (the reference detaches them);
  * layer widths are generic (the reference hard-codes 64 in BatchNorm1d and the dense input);
  * ``GraphConvModel(n_tasks, graph_conv_layers, dense_layer_size, mode)`` (the Keras call
    shape, deepchem/models/graph_models.py:922-933) is accepted as well as the torch one.
Kept as in the torch port: BatchNorm1d(eps=1e-3, momentum=0.99), dropout gated by the
``training`` argument, untrimmed fingerprint output, -1/0 rows for absent molecules, state_dict
keys.
"""
import collections
import logging
import os
import time
from collections.abc import Sequence as SequenceCollection
from typing import List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as _lib_mod
from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .data import NumpyDataset, PackedDataset  # noqa: F401  (re-exported)
from .engine import FlatEngine
from .layers import GraphConv, GraphGather, GraphPool, gemm_mode_code
from .mol_graphs import BatchLayout, pack_convmols
from .synthetic import LazyTake, PackedMols

logger = logging.getLogger(__name__)


def _default_host_workers():
    """Layout-builder threads per process: the cores this process may use, shared between the ranks of the
    node (LOCAL_WORLD_SIZE under torchrun), minus the training and prefetch threads; between 1 and 4."""
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 2
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    return max(1, min(4, cores // ranks - 2))


# upload the exact int8 copy of the features when the shard has one (DCGC_FEATURES_I8=0: always fp32)
_USE_I8 = os.environ.get("DCGC_FEATURES_I8", "1") != "0"


def _lower_thread_priority():
    """Layout workers run at a lower priority than the threads that feed the GPU (kernel launches, uploads): with
    one process per GPU and every core busy (8 ranks on 32 cores) a launch thread that waits for a core stalls the
    device, a layout that finishes a little later does not."""
    try:
        import threading
        os.setpriority(os.PRIO_PROCESS, threading.get_native_id(), 10)
    except Exception:
        pass


def _lib_handle():
    return _lib_mod.lib()


def _lib_check(status):
    return _lib_mod.check(status)


def _chunked_h2d(dst, src):
    """dst.copy_(src) from pinned host memory as a train of moderate asynchronous copies on the current stream
    (one C call, GIL released; dcgc_h2d_chunked).  One 30 MB cudaMemcpyAsync running beside the training kernels
    slowed the forward kernels by up to 20 % of a step on B200; the same bytes in DCGC_H2D_CHUNK_MB (default 0.5)
    MiB pieces cost a fraction of that (scripts/interference.py, profiles/r2_interference.md)."""
    from .mol_graphs import _h2d_chunk_bytes
    step = _h2d_chunk_bytes()
    if step <= 0 or not (src.is_contiguous() and dst.is_contiguous() and src.is_pinned() and dst.is_cuda) or \
            src.dtype != dst.dtype or src.numel() != dst.numel():
        dst.copy_(src, non_blocking=True)
        return
    import ctypes
    from . import _lib
    _lib.check(_lib.lib().dcgc_h2d_chunked(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size(), step,
                                           ctypes.c_void_p(torch.cuda.current_stream(dst.device).cuda_stream)))


class _DoneEvent(object):
    """Stand-in for a CUDA event when the loss already lives on the host."""

    def query(self):
        return True

    def synchronize(self):
        pass


class TrimGraphOutput(nn.Module):
    """Trim to the number of real samples (graphconvmodel.py:21-33).  n_samples stays on the
    host, so no device sync happens here."""

    def forward(self, inputs):
        n_samples = int(inputs[1])
        return inputs[0][0:n_samples]


class _GraphConvTorchModel(nn.Module):
    """GraphConv -> BN -> (dropout) -> GraphPool per layer, Dense+ReLU -> BN -> GraphGather(tanh)
    -> head (graphconvmodel.py:77-249; generic widths as in graph_models.py:835-902)."""

    def __init__(self, n_tasks: int, number_input_features: List[int] = None,
                 graph_conv_layers: List[int] = [64, 64], dense_layer_size: int = 128, dropout=0.0,
                 mode: str = "classification", number_atom_features: int = 75, n_classes: int = 2,
                 batch_normalize: bool = True, uncertainty: bool = False, batch_size: int = 100,
                 gemm_mode: str = "tf32x3", sync_batch_norm: bool = False):
        super(_GraphConvTorchModel, self).__init__()
        if mode not in ['classification', 'regression']:
            raise ValueError("mode must be either 'classification' or 'regression'")
        graph_conv_layers = list(graph_conv_layers)
        if number_input_features is None:
            number_input_features = [number_atom_features] + graph_conv_layers[:-1]
        self.n_tasks, self.n_classes, self.mode = n_tasks, n_classes, mode
        self.uncertainty = uncertainty
        self.gemm_mode = gemm_mode_code(gemm_mode)
        if not isinstance(dropout, SequenceCollection):
            dropout = [dropout] * (len(graph_conv_layers) + 1)
        if len(dropout) != len(graph_conv_layers) + 1:
            raise ValueError('Wrong number of dropout probabilities provided')
        if uncertainty:
            if mode != "regression":
                raise ValueError("Uncertainty is only supported in regression mode")
            if any(d == 0.0 for d in dropout):
                raise ValueError('Dropout must be included in every layer to predict uncertainty')

        self.graph_convs = nn.ModuleList([
            GraphConv(layer_size, input_size, activation_fn=F.relu, gemm_mode=gemm_mode)
            for layer_size, input_size in zip(graph_conv_layers, number_input_features)])

        # sync_batch_norm (not in the reference, whose statistics are per process): training statistics over the atoms
        # of every data-parallel rank, parallel.SyncBatchNorm1d — same parameters, buffers and state_dict keys
        self.sync_batch_norm = bool(sync_batch_norm and batch_normalize)
        if self.sync_batch_norm:
            from .parallel import SyncBatchNorm1d as bn_cls
        else:
            bn_cls = nn.BatchNorm1d

        def bn(width):
            return bn_cls(num_features=width, eps=1e-3, momentum=0.99, affine=True,
                          track_running_stats=True) if batch_normalize else nn.Identity()
        self.batch_norms = nn.ModuleList([bn(c) for c in graph_conv_layers] + [bn(dense_layer_size)])
        self.dropouts = nn.ModuleList([nn.Dropout(rate) if rate > 0.0 else nn.Identity() for rate in dropout])
        self.graph_pools = nn.ModuleList([GraphPool() for _ in graph_conv_layers])
        self.dense = nn.Linear(graph_conv_layers[-1], dense_layer_size)
        self.dense_act = F.relu
        self.graph_gather = GraphGather(batch_size=batch_size, activation_fn=torch.tanh)
        self.trim = TrimGraphOutput()
        if self.mode == 'classification':
            self.reshape_dense = nn.Linear(dense_layer_size * 2, n_tasks * n_classes)
        else:
            self.regression_dense = nn.Linear(dense_layer_size * 2, n_tasks)
            if self.uncertainty:
                self.uncertainty_dense = nn.Linear(dense_layer_size * 2, n_tasks)
                self.uncertainty_trim = TrimGraphOutput()

    def forward(self, inputs, training=False):
        atom_features = inputs[0]
        degree_slice = inputs[1]
        membership = inputs[2]
        n_samples = inputs[3]
        deg_adjs = list(inputs[4:])

        in_layer = atom_features
        for i in range(len(self.graph_convs)):
            gc1 = self.graph_convs[i]([in_layer, degree_slice, membership] + deg_adjs)
            gc1 = self.batch_norms[i](gc1)
            if training:
                gc1 = self.dropouts[i](gc1)
            in_layer = self.graph_pools[i]([gc1, degree_slice, membership] + deg_adjs)
        # atom-level Dense + ReLU through the same grouped-GEMM kernel (single group)
        denseact = ops.GroupLinearFn.apply(in_layer, self.dense.weight.t().contiguous(), self.dense.bias,
                                           ACT_RELU, self.gemm_mode)
        denseact = self.batch_norms[-1](denseact)
        if training:
            denseact = self.dropouts[-1](denseact)
        neural_fingerprint = self.graph_gather([denseact, degree_slice, membership] + deg_adjs)
        if self.mode == 'classification':
            logits = torch.reshape(self.reshape_dense(neural_fingerprint), (-1, self.n_tasks, self.n_classes))
            logits = self.trim([logits, n_samples])
            output = F.softmax(logits, dim=2)
            return [output, logits, neural_fingerprint]
        output = self.trim([self.regression_dense(neural_fingerprint), n_samples])
        if self.uncertainty:
            log_var = self.uncertainty_trim([self.uncertainty_dense(neural_fingerprint), n_samples])
            return [output, torch.exp(log_var), output, log_var, neural_fingerprint]
        return [output, neural_fingerprint]


def _standard_loss(mode, uncertainty):
    """(loss * w).mean() over all elements, as _StandardLoss (torch_model.py:1275-1294) with
    L2Loss (losses.py:76-94) / SoftmaxCrossEntropy (losses.py:236-259) / the uncertainty loss
    (graphconvmodel.py:360-372)."""
    def loss(outputs, labels, weights):
        y, w = labels[0], weights[0]
        if mode == "classification":
            per = -torch.sum(y * F.log_softmax(outputs[0], dim=-1), dim=-1)
        elif uncertainty:
            out, log_var = outputs[0], outputs[1]
            per = torch.square(out - y.reshape(out.shape)) / torch.exp(log_var) + log_var
        else:
            per = F.mse_loss(outputs[0], y.reshape(outputs[0].shape), reduction='none')
        if w.dim() < per.dim():
            w = w.reshape(tuple(w.shape) + (1,) * (per.dim() - w.dim()))
        return (per * w).mean()
    return loss


def _undo_transforms(y, transformers):
    """deepchem.trans.undo_transforms applied as TorchModel._predict does (torch_model.py:625-634): y-transformers are
    undone in reverse order; several outputs cannot be untransformed."""
    if not transformers:
        return y
    if isinstance(y, list):
        if len(y) > 1:
            raise ValueError("predict() does not support Transformers for models with multiple outputs.")
        return [_undo_transforms(v, transformers) for v in y]
    for t in reversed(list(transformers)):
        if getattr(t, "transform_y", False):
            y = t.untransform(y)
    return y


def evaluate_model(model, dataset, metrics, transformers=[], per_task_metrics=False, use_sample_weights=False,
                   n_classes=2):
    """Evaluator.compute_model_performance (deepchem/utils/evaluate.py:246-333): the y-transformers are undone on the
    dataset's labels AND on the predictions before the metrics see them (:303-307), metrics may be
    ``dc.metrics.Metric``-like objects (``compute_metric(y, y_pred, w, per_task_metrics=, n_tasks=, n_classes=,
    use_sample_weights=)``, name in ``.name``) or plain functions f(y_true, y_pred), which get the reference's default
    wrapping: the mean over tasks of f on each task column.  -> {name: score}, and with ``per_task_metrics`` a second
    dict {name: per-task scores}."""
    import inspect
    y = _undo_transforms(np.asarray(dataset.y), transformers)
    w = dataset.w
    y_pred = model.predict(dataset, transformers)
    n_tasks = int(y.shape[1]) if np.ndim(y) > 1 else 1
    if not isinstance(metrics, (list, tuple)):
        metrics = [metrics]
    scores, per_task = {}, {}
    for m in metrics:
        if hasattr(m, "compute_metric"):
            name = getattr(m, "name", m.__class__.__name__)
            try:
                accepted = set(inspect.signature(m.compute_metric).parameters)
            except (TypeError, ValueError):
                accepted = set()
            kw = {k: v for k, v in (("per_task_metrics", per_task_metrics), ("n_tasks", n_tasks),
                                    ("n_classes", n_classes), ("use_sample_weights", use_sample_weights))
                  if k in accepted}
            res = m.compute_metric(y, y_pred, w, **kw)
            if per_task_metrics and "per_task_metrics" in kw:
                scores[name], per_task[name] = res
            else:
                scores[name] = res
        else:
            name = getattr(m, "__name__", "metric")
            yt = np.reshape(y, (len(y), n_tasks))
            yp = np.reshape(y_pred, (len(y_pred), n_tasks, -1))
            yp = yp[..., 0] if yp.shape[-1] == 1 else yp
            per = [float(m(yt[:, t], yp[:, t])) for t in range(n_tasks)]
            scores[name] = float(np.mean(per))
            per_task[name] = per
    return (scores, per_task) if per_task_metrics else scores


def to_one_hot(y, n_classes=2):
    """deepchem.metrics.to_one_hot"""
    y = np.asarray(y).astype(np.int64).reshape(-1)
    out = np.zeros((y.shape[0], n_classes), dtype=np.float32)
    out[np.arange(y.shape[0]), y] = 1
    return out


class BatchInputs(list):
    """The list default_generator yields ([features, deg_slice, membership, n_samples,
    deg_adj_1..10], numpy views of the layout slab) carrying the BatchLayout itself so that
    _prepare_batch can move everything with one H2D copy."""
    layout = None
    packed_features = None
    packed_features_pinned = None
    packed_features_i8_pinned = None


class _DeviceSlot(object):
    """Reusable device + pinned-host buffers of one in-flight batch (grow-only).  The steady-state step then
    makes no cudaMalloc / cudaHostAlloc call: with per-batch tensors of slightly different sizes the caching
    allocators occasionally had to go to the driver, which synchronises the device (seen as multi-millisecond
    stalls of the end-to-end step, worst with several ranks on one host)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.free_event = None      # recorded on the consumer's stream when it is done with this slot
        self.copied_event = None    # recorded on the producer's stream after the H2D copies of this slot

    def get(self, name, numel, dtype, pinned=False):
        t = self.bufs.get(name)
        if t is None or t.numel() < numel:
            cap = int(numel * 1.15) + 1024
            t = torch.empty(cap, dtype=dtype, pin_memory=True) if pinned else \
                torch.empty(cap, dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t


_PINNED_DIRECT_LIMIT = 192 << 20    # predict(): largest result copied straight into ONE page-locked array; a larger one
                                    # would spend longer in cudaHostAlloc (~1 GB/s) than the GPU needs for the whole pass
                                    # (measured: 1.28 GB of PCBA probabilities, 1.10 s per 1.25 M molecules against 0.28 s
                                    # of device work) and goes through a small reusable staging ring instead


class _StagedResult(object):
    """Large ``predict`` results: device outputs are copied asynchronously into a small ring of reusable page-locked
    buffers (allocated once per model) and two helper threads move every filled buffer into ordinary (pageable) result
    arrays while the GPU computes the next batches.  The copies out of a buffer start when the CUDA event recorded after
    its last device-to-host copy has fired; a buffer is reused when its copies are done."""

    BUF_BYTES = 16 << 20
    N_BUF = 4

    def __init__(self, model, shapes):
        import queue
        import threading
        self.results = [np.empty(sh, dtype=np.float32) for sh in shapes]
        self.offs = [0] * len(shapes)
        ring = getattr(model, "_result_ring", None)
        need = max(self.BUF_BYTES, max(int(np.prod(sh[1:])) for sh in shapes) * 4 * max(1, int(model.batch_size)))
        if ring is None or ring[0].numel() < need:
            ring = [torch.empty(need, dtype=torch.uint8, pin_memory=True) for _ in range(self.N_BUF)]
            model._result_ring = ring
        self.ring = ring
        self.free = queue.Queue()
        for b in ring:
            self.free.put(b)
        self.todo = queue.Queue()
        self.error = None
        self.cur, self.cur_off, self.cur_recs = self.free.get(), 0, []
        self.threads = [threading.Thread(target=self._copier, daemon=True) for _ in range(2)]
        for t in self.threads:
            t.start()

    def _copier(self):
        while True:
            item = self.todo.get()
            if item is None:
                return
            buf, ev, recs = item
            try:
                ev.synchronize()
                host = buf.numpy()
                for off, nbytes, i, row0, nrows in recs:
                    dst = self.results[i][row0:row0 + nrows]
                    np.copyto(dst.reshape(-1), host[off:off + nbytes].view(np.float32))
            except BaseException as e:      # surfaced in finish()
                self.error = e
            finally:
                self.free.put(buf)

    def _seal(self):
        if self.cur_recs:
            # blocking=True: a host thread that waits on this event SLEEPS (cudaEventBlockingSync) instead of spinning on
            # a core — with 4 cores per rank (8 ranks on a 32-core host) the spinning copier / prefetch threads starved
            # the thread that launches the kernels (8 ms to issue one forward pass, profiles/r5w_host_threads.md)
            ev = torch.cuda.Event(blocking=True)
            ev.record()
            self.todo.put((self.cur, ev, self.cur_recs))
            self.cur, self.cur_off, self.cur_recs = self.free.get(), 0, []

    def add(self, vals):
        for i, v in enumerate(vals):
            nrows, nbytes = v.shape[0], v.numel() * 4
            if self.offs[i] + nrows > self.results[i].shape[0]:
                raise ValueError("predict: more output rows than the dataset announced")
            if self.cur_off + nbytes > self.cur.numel():
                self._seal()
            dst = self.cur[self.cur_off:self.cur_off + nbytes].view(torch.float32).view(v.shape)
            dst.copy_(v, non_blocking=True)
            self.cur_recs.append((self.cur_off, nbytes, i, self.offs[i], nrows))
            self.cur_off += (nbytes + 255) // 256 * 256
            self.offs[i] += nrows

    def finish(self):
        self._seal()
        for _ in self.threads:
            self.todo.put(None)
        for t in self.threads:
            t.join()
        if self.error is not None:
            raise self.error
        return [r[:o] for r, o in zip(self.results, self.offs)]


class _PinnedRing(object):
    """Reusable page-locked staging buffers for the host side of a batch (page-locked allocations cost milliseconds,
    so they are made once).  A slot is handed out again only when (1) nothing references the arrays built in it any
    more — a weak reference to the root numpy array every view of the batch hangs off (or to the tensor handed to the
    gather) — and (2) the upload that read it has finished.  A caller that keeps batches alive — ``list(model.
    default_generator(ds))``, multi-epoch reuse of a materialised list, a consumer lagging behind — therefore never
    sees an earlier batch overwritten (the reference's generator yields independent arrays, graphconvmodel.py:382-422):
    the ring grows up to ``max_slots`` and, past that, ``take`` returns ``(None, None)`` and the caller builds into
    ordinary memory."""

    def __init__(self, max_slots=64, pin=True):
        import threading
        self.pin = pin
        self.slots = []              # [tensor, upload event or None, weakref to the owner or None / _BUSY]
        self.next = 0
        self.max_slots = max_slots
        self.gc_at = 12              # ring length from which a full ring first tries the cyclic collector
        self.lock = threading.Lock()

    _BUSY = object()                 # handed out, owner not registered yet

    @staticmethod
    def _free(slot):
        o = slot[2]
        return o is None or (o is not _PinnedRing._BUSY and o() is None)

    def take(self, nbytes):
        import weakref
        slot = None
        with self.lock:
            n = len(self.slots)
            for attempt in range(2):
                for k in range(n):
                    cand = self.slots[(self.next + k) % n]
                    if self._free(cand):
                        slot = cand
                        self.next = (self.next + k + 1) % n
                        break
                if slot is not None or n < self.gc_at or attempt:
                    break
                import gc
                gc.collect()         # batches kept alive only by reference cycles: cheaper than a page-locked allocation
            if slot is None:
                if n >= self.max_slots:
                    return None, None
                slot = [torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=self.pin), None, None]
                self.slots.append(slot)
            slot[2] = self._BUSY
        if slot[1] is not None:
            slot[1].synchronize()    # the H2D copy that read the previous batch out of this buffer
            slot[1] = None
        if slot[0].numel() < nbytes:
            slot[0] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=self.pin)
        root = slot[0].numpy()
        slot[2] = weakref.ref(root)
        return slot, root

    @staticmethod
    def own(slot, obj):
        """The object whose lifetime keeps the slot busy (default: the root array returned by take)."""
        import weakref
        slot[2] = weakref.ref(obj)


class _Prefetcher(object):
    """Runs ``model._prepare_batch`` for upcoming batches on a helper thread and a side CUDA stream;
    the consumer's stream waits on an event before touching a prepared batch."""

    def __init__(self, model, generator, depth=2):
        import queue
        import threading
        self.model, self.generator = model, generator
        self.q = queue.Queue(maxsize=max(1, depth))
        # one side stream per model: the caching allocator pools blocks per stream, so a fresh
        # stream per fit() call would pay cudaMalloc again for every staging tensor
        if model._prefetch_stream is None:
            model._prefetch_stream = torch.cuda.Stream(device=model.device)
        self.stream = model._prefetch_stream
        n_slots = max(1, depth) + 2          # queued + being consumed + being prepared
        while len(model._device_slots) < n_slots:
            model._device_slots.append(_DeviceSlot(model.device))
        self.slots = model._device_slots[:n_slots]
        self.error = None
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            torch.cuda.set_device(self.model.device)
            k = 0
            tr = self.model._pipe_trace          # optional stage timers (DCGC_PIPE_TRACE=1), seconds
            clock = time.perf_counter
            it = iter(self.generator)
            while True:
                t0 = clock()
                try:
                    batch = next(it)
                except StopIteration:
                    break
                if self.stop:
                    break
                t1 = clock()
                slot = self.slots[k % len(self.slots)]
                k += 1
                with torch.cuda.stream(self.stream):
                    if slot.free_event is not None:
                        self.stream.wait_event(slot.free_event)      # the consumer finished with these buffers
                    if slot.copied_event is not None:
                        slot.copied_event.synchronize()              # pinned staging may be rewritten
                    t2 = clock()
                    fe = self.model._last_fwd_event
                    if fe is not None:
                        self.stream.wait_event(fe)                   # upload beside the backward pass
                    prepared = self.model._prepare_batch(batch, slot)
                    ev = torch.cuda.Event(blocking=True)
                    ev.record(self.stream)
                    slot.copied_event = ev
                t3 = clock()
                self.q.put((prepared, ev, slot))
                if tr is not None:
                    t4 = clock()
                    tr["pf_wait_generator"] += t1 - t0
                    tr["pf_wait_slot"] += t2 - t1
                    tr["pf_prepare"] += t3 - t2
                    tr["pf_wait_queue"] += t4 - t3
                    tr["pf_batches"] += 1
        except BaseException as e:      # surfaced in the consumer
            self.error = e
        finally:
            self.q.put(None)

    def __iter__(self):
        try:
            prev = None
            while True:
                if prev is not None:
                    # the consumer has issued all its work on the previous batch: its slot may be refilled once
                    # that work has run.  This must be published BEFORE q.get(): taking an item is what unblocks a
                    # producer waiting on the full queue, and its next slot is this one — set afterwards, the
                    # producer could read the slot's previous (long completed) event and overwrite device buffers
                    # that the kernels of this batch were still reading (seen as an illegal memory access with 8
                    # distinct batches per epoch).
                    fe = torch.cuda.Event()
                    fe.record(torch.cuda.current_stream())
                    prev.free_event = fe
                    prev = None
                item = self.q.get()
                if item is None:
                    break
                prepared, ev, slot = item
                main = torch.cuda.current_stream()
                main.wait_event(ev)
                prev = slot
                yield prepared
            if self.error is not None:
                raise self.error
        finally:
            self.stop = True
            while self.thread.is_alive():       # unblock a producer stuck on a full queue
                try:
                    self.q.get_nowait()
                except Exception:
                    pass
                self.thread.join(timeout=0.05)


def _check_f16_range(model, any_path=False):
    """The fused engine runs its forward GEMMs with fp16 operand halves (DCGC_GEMM_F16X3, csrc/gemm_tc.cu): raise if a
    forward GEMM has seen an operand outside fp16's range — its results are then wrong, not merely inexact.
    ``any_path``: the model's per-layer path uses them too (D-MPNN)."""
    if (model._engine is None and not any_path) or model.device.type != "cuda":
        return
    if _lib_mod.lib().dcgc_tc_f16_overflow() == 1:
        raise FloatingPointError("an activation or weight above 3750 reached a forward GEMM that runs with fp16 operand "
                                 "halves (operands are scaled by 16 before the split); set DCGC_FWD_F16X3=0 (TF32 "
                                 "halves) for data of this range")


class GraphConvModel(object):
    """Graph convolutional model on the B200 path.

    ``GraphConvModel(n_tasks, number_input_features=[75, 64], graph_conv_layers=[64, 64], ...)``
    (torch reference, graphconvmodel.py:284-296) and ``GraphConvModel(n_tasks, graph_conv_layers,
    dense_layer_size, ...)`` (Keras reference, graph_models.py:922-933) are both accepted; a single
    positional list is read as ``number_input_features`` only if it starts with
    ``number_atom_features``.

    ``sync_batch_norm=True`` (an addition): in a data-parallel run the BatchNorm layers take their training
    statistics over the atoms of every rank (``parallel.SyncBatchNorm1d``: two small all-reduces per layer and
    direction) instead of per process as the reference does; the model then runs on the per-layer autograd path, not
    the fused engine.  Checkpoints are interchangeable with the default model.
    """

    def __init__(self, n_tasks, *args, graph_conv_layers=None, number_input_features=None,
                 dense_layer_size=128, dropout=0.0, mode="classification", number_atom_features=75,
                 n_classes=2, batch_size=100, batch_normalize=True, uncertainty=False,
                 learning_rate=0.001, model_dir=None, device=None, gemm_mode="tf32x3", log_frequency=100,
                 use_engine=True, sync_batch_norm=False, **kwargs):
        args = list(args)
        if args and isinstance(args[0], (list, tuple)):
            first = list(args.pop(0))
            if args and isinstance(args[0], (list, tuple)):      # torch: (n_tasks, in_widths, conv_widths)
                number_input_features, graph_conv_layers = first, list(args.pop(0))
            elif graph_conv_layers is not None or (first and first[0] == number_atom_features):
                number_input_features = first                    # torch: (n_tasks, in_widths)
            else:
                graph_conv_layers = first                        # Keras: (n_tasks, conv_widths, dense, ...)
        if args:
            dense_layer_size = args.pop(0)
        if args:
            dropout = args.pop(0)
        if args:
            mode = args.pop(0)
        if args:
            raise TypeError("too many positional arguments")
        if graph_conv_layers is None:
            graph_conv_layers = [64, 64]
        self.mode, self.n_tasks, self.n_classes = mode, n_tasks, n_classes
        self.batch_size, self.uncertainty = batch_size, uncertainty
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("deepchem_b200.GraphConvModel needs a CUDA device (B200); "
                                   "there is no CPU path")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.model = _GraphConvTorchModel(
            n_tasks, graph_conv_layers=graph_conv_layers, number_input_features=number_input_features,
            dense_layer_size=dense_layer_size, dropout=dropout, mode=mode,
            number_atom_features=number_atom_features, n_classes=n_classes,
            batch_normalize=batch_normalize, uncertainty=uncertainty, batch_size=batch_size,
            gemm_mode=gemm_mode, sync_batch_norm=sync_batch_norm).to(self.device)
        if mode == "classification":
            self.output_types = ['prediction', 'loss', 'embedding']
        elif uncertainty:
            self.output_types = ['prediction', 'variance', 'loss', 'loss', 'embedding']
        else:
            self.output_types = ['prediction', 'embedding']
        self._prediction_outputs = [i for i, t in enumerate(self.output_types) if t == 'prediction']
        self._loss_outputs = [i for i, t in enumerate(self.output_types) if t == 'loss'] or self._prediction_outputs
        self._variance_outputs = [i for i, t in enumerate(self.output_types) if t == 'variance']
        self._embedding_outputs = [i for i, t in enumerate(self.output_types) if t == 'embedding']
        self._loss_fn = _standard_loss(mode, uncertainty)
        self.learning_rate = learning_rate
        self._pytorch_optimizer = torch.optim.Adam(self.model.parameters(), lr=learning_rate,
                                                   betas=(0.9, 0.999), eps=1e-8)
        self._global_step = 0
        self._grad_slab = None
        self._dp = False
        self._prefetch_stream = None
        self._loss_ring = [torch.zeros((), dtype=torch.float32).pin_memory() if self.device.type == "cuda"
                           else torch.zeros(()) for _ in range(4)]
        self._device_slots = []     # reusable per-batch device buffers of the prefetch pipeline
        # upload phase (DCGC_H2D_PHASE=bwd, off by default): the prefetch stream starts a batch's uploads only after
        # the forward pass of the newest launched step.  Measured on B200 (profiles/r2_interference.md): no gain over
        # the chunked int8 upload, so it stays an experiment switch.
        self._fwd_events = None
        self._last_fwd_event = None
        self._bn_steps_pending = 0
        self.model.register_state_dict_pre_hook(self._flush_bn_counters)
        if self.device.type == "cuda" and os.environ.get("DCGC_H2D_PHASE", "none") == "bwd":
            self._fwd_events = []
            with torch.cuda.device(self.device):
                for _ in range(8):
                    ev = torch.cuda.Event()
                    ev.record()                  # materialises the cudaEvent_t handle
                    self._fwd_events.append(ev)
        # per-stage host timers of the fit pipeline (scripts/e2e_stages.py); None = off
        self._pipe_trace = collections.defaultdict(float) if os.environ.get("DCGC_PIPE_TRACE") == "1" else None
        self._staging = _PinnedRing()          # reusable pinned slabs of the batch layouts
        self._feat_staging = _PinnedRing()     # same, for the gathered features of shuffled batches
        # host threads building batch layouts ahead of the GPU (default_generator)
        self.host_workers = int(os.environ.get("DCGC_HOST_WORKERS", _default_host_workers()))
        # fused whole-model engine (flat parameter slab, one C call per step) when the model shape
        # allows it; otherwise the per-layer autograd ops are used.
        self._engine = None
        if use_engine and FlatEngine.eligible(self.model):
            self._engine = FlatEngine(self.model, self.device, lr=learning_rate)
        self.log_frequency = log_frequency
        self.model_dir = model_dir
        self.number_atom_features = number_atom_features

    # ------------------------------------------------------------------ batching
    def default_generator(self, dataset, epochs=1, mode='fit', deterministic=True, pad_batches=True,
                          workers=None):
        """Dataset -> (inputs, [y], [w]) per batch (graphconvmodel.py:382-422).  The layout comes
        from the C++ builder instead of ConvMol.agglomerate_mols.  ``workers`` > 1 builds the layouts
        of upcoming batches concurrently on a thread pool (the C builder releases the GIL); batches
        are still yielded in dataset order."""
        if workers is None:
            workers = self.host_workers

        def batches():
            kw = {}
            if isinstance(dataset, PackedDataset) and os.environ.get("DCGC_LAZY_TAKE", "1") != "0":
                kw["lazy"] = True          # shuffled / padded batches are gathered by the layout workers (batch_inputs)
            for (X_b, y_b, w_b, ids_b) in dataset.iterbatches(batch_size=self.batch_size, epochs=epochs,
                                                              deterministic=deterministic,
                                                              pad_batches=pad_batches, **kw):
                if y_b is not None and self.mode == 'classification' and not (mode == 'predict'):
                    y_b = to_one_hot(np.asarray(y_b).flatten(), self.n_classes).reshape(
                        -1, self.n_tasks, self.n_classes)
                yield X_b, y_b, w_b

        if workers <= 1:
            for X_b, y_b, w_b in batches():
                yield (self.batch_inputs(X_b), [y_b], [w_b])
            return
        import collections
        from concurrent.futures import ThreadPoolExecutor
        pending = collections.deque()
        with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="dcgc-layout",
                                initializer=_lower_thread_priority) as pool:
            for X_b, y_b, w_b in batches():
                pending.append((pool.submit(self.batch_inputs, X_b), y_b, w_b))
                if len(pending) > workers:
                    fut, y0, w0 = pending.popleft()
                    yield (fut.result(), [y0], [w0])
            while pending:
                fut, y0, w0 = pending.popleft()
                yield (fut.result(), [y0], [w0])

    def _staging_slab(self, nbytes):
        """-> (slot, root) from the ring of pinned slabs, or (None, None) when every slot is still referenced."""
        return self._staging.take(nbytes)

    def _feature_staging(self, nbytes):
        """Same ring discipline for the gathered features of a shuffled batch."""
        return self._feat_staging.take(nbytes)

    def batch_inputs(self, X_b, pinned=True):
        pinned = pinned and torch.cuda.is_available()     # host-only callers (tests) get plain memory
        feat_slots = []
        if isinstance(X_b, LazyTake):
            # a shuffled batch: gather the molecules here (worker thread, GIL released in C), features straight into
            # page-locked staging memory — their exact int8 copy alone when the shard has one
            alloc = None
            if pinned:
                def alloc(n_atoms, n_feat, dtype):
                    nbytes = n_atoms * n_feat * np.dtype(dtype).itemsize
                    tdt = torch.int8 if np.dtype(dtype) == np.int8 else torch.float32
                    fs, _ = self._feature_staging(nbytes)
                    if fs is None:               # every slot still referenced by batches the caller keeps
                        return torch.empty((n_atoms, n_feat), dtype=tdt)
                    t = fs[0][:nbytes].view(tdt).view(n_atoms, n_feat)
                    _PinnedRing.own(fs, t)       # busy while the gathered shard (its _pin / _pin_i8) is alive
                    feat_slots.append(fs)
                    return t
            X_b = X_b.resolve(alloc=alloc, prefer_i8=bool(pinned and _USE_I8), n_threads=1)
        packed = X_b if isinstance(X_b, PackedMols) else pack_convmols(X_b)
        n_seg = max(self.batch_size, packed.n_mols)
        slot = root = None
        if pinned:
            # upper bound of the slab size without running the planner: 11 int32 arrays over atoms /
            # edges / segments plus alignment
            need = 4 * (5 * (packed.n_atoms + 2) + 3 * int(packed.adj_ptr[-1]) + (n_seg + 2)
                        + 4 * (packed.n_atoms // 128 + 12) + 12 * (packed.n_atoms // 4 + 4)) + 256 * 16
            slot, root = self._staging_slab(need)
        layout = BatchLayout.build(packed, n_segments=n_seg, pinned=pinned and slot is not None,
                                   staging=slot[0] if slot is not None else None,
                                   staging_root=root if slot is not None else None)
        layout._staging_slot = slot
        inputs = BatchInputs([None, layout.deg_slice, layout.membership, np.array(packed.n_mols)]
                             + layout.deg_adjacency_lists()[1:])
        inputs.layout = layout
        inputs.packed_features = packed.features
        # page-locked torch view of the same rows when the shard was pinned: the H2D copy is then a true
        # asynchronous DMA.  (A numpy view of pinned memory goes through torch's pageable path: a blocking
        # cudaMemcpy that holds the driver's context lock and stalls the training thread's kernel launches —
        # measured: the 0.8 ms copy serialised with the 1.7 ms step instead of overlapping it.)
        pin = getattr(packed, "_pin", None)
        inputs.packed_features_pinned = pin if (pin is not None and tuple(pin.shape) == packed.features.shape) else None
        # exact int8 copy of the same rows (PackedMols.compact): a quarter of the upload
        p8 = getattr(packed, "_pin_i8", None)
        inputs.packed_features_i8_pinned = p8 if (p8 is not None and tuple(p8.shape) == packed.features.shape) else None
        inputs.feature_staging_slots = feat_slots
        return inputs

    def _prepare_batch(self, batch, slot=None):
        """Host -> device boundary (torch_model.py:923-952): one copy for the integer slab, one
        for the (unpermuted) features, then a device-side row permutation into degree-major
        order with rows padded to a 16-byte multiple.  ``slot`` (prefetch pipeline): reusable device /
        pinned buffers to build into instead of fresh allocations."""
        inputs, labels, weights = batch
        if getattr(inputs, "layout", None) is None:
            raise TypeError("inputs must come from GraphConvModel.default_generator / batch_inputs")
        layout = inputs.layout
        buf = slot.get("slab", int(layout.info.slab_bytes), torch.uint8) if slot is not None else None
        rec = slot.get("mgrec", 80 * layout.n_atoms, torch.uint8) if slot is not None else None
        topo = layout.to_device(self.device, buffer=buf, record_buffer=rec)
        sslot = getattr(layout, "_staging_slot", None)
        if sslot is not None:
            sslot[1] = torch.cuda.Event(blocking=True)
            sslot[1].record(torch.cuda.current_stream())
        feats = getattr(inputs, "packed_features_i8_pinned", None) if (slot is not None and _USE_I8) else None
        if feats is None:
            feats = getattr(inputs, "packed_features_pinned", None)
        if feats is None:
            feats = torch.from_numpy(np.ascontiguousarray(inputs.packed_features, dtype=np.float32))
        n, f = feats.shape
        if slot is not None:
            fdev = slot.get("feats8" if feats.dtype == torch.int8 else "feats", n * f, feats.dtype)[:n * f].view(n, f)
            _chunked_h2d(fdev, feats)
            x = ops.permute_rows(fdev, topo.perm, out=slot.get("x", n * ((f + 3) // 4 * 4), torch.float32))
        else:
            x = ops.permute_rows(feats.to(self.device, non_blocking=True), topo.perm)
        fslots = getattr(inputs, "feature_staging_slots", None)
        if fslots:                             # the gathered features may be overwritten once this upload is done
            fev = torch.cuda.Event(blocking=True)
            fev.record(torch.cuda.current_stream())
            for fs in fslots:
                fs[1] = fev
        x._dcgc_zero_padded = True
        # integer-valued features (the shard keeps their exact int8 copy, PackedMols.compact): the first layer's GEMM
        # operands are exact in tf32 and the engine skips the identically-zero lo(A) term (dcgc_gcmodel_config.input_exact)
        x._dcgc_input_exact = bool(feats.dtype == torch.int8
                                   or getattr(inputs, "packed_features_i8_pinned", None) is not None)
        dev_inputs = topo.model_inputs(x, n_samples=int(inputs[3]))

        def conv(arrs, tag):
            out = []
            for i, a in enumerate(arrs or []):
                if a is None:
                    out.append(None)
                    continue
                a = np.asarray(a)
                if a.dtype == np.float64:
                    a = a.astype(np.float32)
                a = np.ascontiguousarray(a)
                t = torch.from_numpy(a)
                if slot is not None and self.device.type == "cuda":
                    host = slot.get("h%s%d" % (tag, i), t.numel(), t.dtype, pinned=True)[:t.numel()].view(t.shape)
                    host.copy_(t)
                    dev = slot.get("d%s%d" % (tag, i), t.numel(), t.dtype)[:t.numel()].view(t.shape)
                    dev.copy_(host, non_blocking=True)
                    out.append(dev)
                else:
                    if self.device.type == "cuda":
                        t = t.pin_memory()
                    out.append(t.to(self.device, non_blocking=True))
            return out
        return dev_inputs, conv(labels, "y"), conv(weights, "w")

    # ------------------------------------------------------------------ training
    def fit(self, dataset, nb_epoch=10, max_checkpoints_to_keep=5, checkpoint_interval=1000,
            deterministic=False, restore=False, callbacks=[], all_losses=None):
        return self.fit_generator(
            self.default_generator(dataset, epochs=nb_epoch, deterministic=deterministic),
            max_checkpoints_to_keep, checkpoint_interval, restore, callbacks=callbacks,
            all_losses=all_losses)

    def fit_generator(self, generator, max_checkpoints_to_keep=5, checkpoint_interval=1000, restore=False,
                      callbacks=[], all_losses=None, prefetch=2):
        """Train on (inputs, labels, weights) batches (torch_model.py:345-496).  With ``prefetch`` > 0
        the host side of the next batches (C++ layout build, H2D copies on a side stream, device
        permutation) runs in a helper thread while the GPU works on the current one — the role the
        reference gives to DiskDataset's shard-prefetch thread (data/datasets.py:1670-1693)."""
        if not isinstance(callbacks, SequenceCollection):
            callbacks = [callbacks]
        self.model.train()
        t0 = time.time()
        # Every step's loss is read back to the host (4 bytes, pinned, asynchronous) but consumed one step late:
        # the device never waits for the Python loop between steps (a blocking float(loss) per step cost 0.3 ms
        # of a 2 ms step).  Logging / all_losses follow the reference's cadence (torch_model.py:453-463).
        pending = collections.deque()           # (step, pinned scalar, event)
        state = {"sum": 0.0, "n": 0, "last": 0.0}

        def consume(limit):
            # take the losses whose copies have landed; block only while more than `limit` are outstanding
            while pending and (len(pending) > limit or pending[0][2].query()):
                step_i, host, ev = pending.popleft()
                ev.synchronize()
                state["sum"] += float(host)
                state["n"] += 1
                if step_i % self.log_frequency == 0:
                    state["last"] = state["sum"] / state["n"]
                    logger.info('Ending global_step %d: Average loss %g' % (step_i, state["last"]))
                    if all_losses is not None:
                        all_losses.append(state["last"])
                    state["sum"], state["n"] = 0.0, 0

        prepared_iter = _Prefetcher(self, generator, prefetch) if prefetch else \
            (self._prepare_batch(b) for b in generator)
        self._active_prefetcher = prepared_iter if prefetch else None      # (bench.py reads its queue depth)
        tr = self._pipe_trace
        t_prev = time.perf_counter()
        for prepared in prepared_iter:
            if restore:
                self.restore()
                restore = False
            if tr is not None:
                t_a = time.perf_counter()
                tr["fit_wait_batch"] += t_a - t_prev
            batch_loss = self._train_step(*prepared)
            if tr is not None:
                t_prev = time.perf_counter()
                tr["fit_train_step_host"] += t_prev - t_a
                tr["fit_steps"] += 1
            self._global_step += 1
            step = self._global_step
            if batch_loss.is_cuda:
                host = self._loss_ring[step % len(self._loss_ring)]
                host.copy_(batch_loss.detach(), non_blocking=True)
                ev = torch.cuda.Event(blocking=True)
                ev.record()
            else:
                host, ev = batch_loss.detach().clone(), _DoneEvent()
            pending.append((step, host, ev))
            consume(len(self._loss_ring) - 2)
            if self.model_dir and checkpoint_interval > 0 and step % checkpoint_interval == checkpoint_interval - 1:
                self.save_checkpoint(max_checkpoints_to_keep)
            for c in callbacks:
                try:
                    c(self, step, iteration_loss=batch_loss)
                except TypeError:
                    c(self, step)
        consume(0)
        while pending:
            consume(-1)
        if state["n"] > 0:
            state["last"] = state["sum"] / state["n"]
            if all_losses is not None:
                all_losses.append(state["last"])
        if self.model_dir and checkpoint_interval > 0:
            self.save_checkpoint(max_checkpoints_to_keep)
        logger.info("TIMING: model fitting took %0.3f s" % (time.time() - t0))
        self._flush_bn_counters()
        _check_f16_range(self)
        return state["last"]

    def _train_step(self, inputs, labels, weights):
        """zero_grad, forward, loss, backward, Adam step (torch_model.py:435-443).  The reference
        never passes training=True here, so dropout stays off during fit."""
        if self._engine is not None:
            return self._engine_step(inputs, labels, weights)
        slab = self._grad_slab
        if slab is None:
            self._pytorch_optimizer.zero_grad(set_to_none=True)
        else:
            slab.zero()
            slab.attach()
        outputs = self.model(inputs)
        outputs = [outputs[i] for i in self._loss_outputs]
        loss = self._loss_fn(outputs, labels, weights)
        loss.backward()
        if slab is not None:
            # data parallel: one all-reduce of the flat gradient slab, then identical Adam steps
            slab.collect()
            slab.all_reduce_mean()
        self._pytorch_optimizer.step()
        return loss

    def _engine_step(self, inputs, labels, weights):
        """One C call for forward + loss + backward over the flat slabs, one all-reduce of the
        gradient slab in data parallel, one fused Adam launch."""
        eng = self._engine
        topo = inputs[1]._dcgc_topology
        w = weights[0] if weights and weights[0] is not None else None
        fe = None
        if self._fwd_events is not None:
            # the C call records this event between forward and backward; the prefetch stream waits on the latest
            # one before it uploads a batch, so uploads run beside the backward pass (see _Prefetcher._run)
            fe = self._fwd_events[self._global_step % len(self._fwd_events)]
        dp = False
        if self._dp:
            from .parallel import world_size
            dp = world_size() > 1
        # Per-slice exchange beside the backward pass, or ONE all-reduce after the step?  Measured: device-timed steps of
        # 1.2151 against 1.2246 ms at 2 GPUs for the overlapped form and 1.128 against 1.110 at 8 (profiles/r5v_dp8.md:
        # NCCL's CTAs take SMs away from the persistent one-CTA-per-SM kernels of the backward pass), but END TO END the
        # four asynchronous collective calls cost the launching thread 0.62 ms per step against 0.15 for one call
        # (scripts/e2e_trace_dp.py, 2 GPUs: 1.190 against 1.048 ms per step).  Default: one all-reduce;
        # DCGC_OVERLAP=1 selects the per-slice form.
        if os.environ.get("DCGC_NO_OVERLAP", "0") == "1":
            overlap = False
        elif os.environ.get("DCGC_OVERLAP", "0") == "1":
            overlap = dp
        else:
            overlap = False
        gev = self._dp_events() if overlap else None
        tr = self._pipe_trace
        t_0 = time.perf_counter() if tr is not None else 0.0
        loss = eng.train_step(topo, inputs[0], labels[0].contiguous(), w.contiguous() if w is not None else None,
                              int(inputs[3]), forward_event=fe, grad_events=gev)
        t_1 = time.perf_counter() if tr is not None else 0.0
        if fe is not None:
            self._last_fwd_event = fe
        scale = 1.0
        if dp and not overlap:
            import torch.distributed as dist
            from .parallel import world_size
            dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM)
            scale = 1.0 / world_size()
        elif dp:
            # Overlapped gradient exchange: the C call above only ENQUEUED the step, and recorded one event
            # per gradient slice in the order the backward pass finishes them (head + dense, then conv layer L-1 .. 0).
            # Each slice is all-reduced on a communication stream as soon as its event fires, beside the rest of the
            # backward pass; the Adam launch waits for the last one.  (One all-reduce after the whole step cost 77 us of
            # a 1.23 ms step at 8 GPUs.)
            import torch.distributed as dist
            from .parallel import world_size
            comm = self._comm_stream
            works = []
            for ev, (lo, hi) in zip(gev, eng.grad_slices()):
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    works.append(dist.all_reduce(eng.grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            for wk in works:
                wk.wait()                    # stream-level: the training stream waits, the host does not
            scale = 1.0 / world_size()
        t_2 = time.perf_counter() if tr is not None else 0.0
        eng.adam_step(scale)
        t_3 = time.perf_counter() if tr is not None else 0.0
        if eng.cfg.batch_norm:
            # BatchNorm1d.num_batches_tracked: counted on the host, added to the buffers when somebody can see them
            # (state_dict / the end of fit_generator) — a launch and 50 us of the launching thread per step otherwise
            self._bn_steps_pending += 1
        if tr is not None:
            t_4 = time.perf_counter()
            tr["step_c_call"] += t_1 - t_0
            tr["step_allreduce"] += t_2 - t_1
            tr["step_adam"] += t_3 - t_2
            tr["step_bn_counters"] += t_4 - t_3
        return loss

    def _flush_bn_counters(self, *unused):
        n = self._bn_steps_pending
        if n:
            self._bn_steps_pending = 0
            torch._foreach_add_([bn.num_batches_tracked for bn in self.model.batch_norms], n)

    def _dp_events(self):
        """Events of the gradient slices (one set per step in flight) and the communication stream, made on first use."""
        if getattr(self, "_grad_event_sets", None) is None:
            n = len(self._engine.grad_slices())
            self._grad_event_sets = []
            with torch.cuda.device(self.device):
                for _ in range(4):
                    evs = [torch.cuda.Event() for _ in range(n)]
                    for e in evs:
                        e.record()           # materialises the cudaEvent_t handle
                    self._grad_event_sets.append(evs)
                self._comm_stream = torch.cuda.Stream(device=self.device)
        return self._grad_event_sets[self._global_step % len(self._grad_event_sets)]

    def enable_data_parallel(self):
        """Average gradients over the default process group every step (equal per-rank batches
        and a mean loss make the averaged gradient exact).  Parameters are broadcast
        from rank 0 first so that every replica starts identical."""
        import torch.distributed as dist
        from .parallel import GradSlab, world_size
        self._dp = True
        if self._engine is not None:
            if world_size() > 1:
                dist.broadcast(self._engine.params, src=0)
                dist.broadcast(self._engine.bn_running, src=0)
            return self
        if world_size() > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=0)
        self._grad_slab = GradSlab(self.model.parameters())
        return self

    def fit_on_batch(self, X, y, w):
        """One training step on one batch of ConvMol-like molecules (or a PackedMols)."""
        self.model.train()
        n = len(X)
        ds = PackedDataset(X, y, w) if isinstance(X, PackedMols) else NumpyDataset(X, y, w)
        bs, self.batch_size = self.batch_size, max(self.batch_size, n)
        try:
            batch = next(self.default_generator(ds, deterministic=True, pad_batches=False))
        finally:
            self.batch_size = bs
        loss = self._train_step(*self._prepare_batch(batch))
        self._global_step += 1
        self._flush_bn_counters()
        return float(loss)

    # ------------------------------------------------------------------ inference
    def _predict(self, generator, output_idx, n_rows=None):
        """Forward-only pass over the batches (torch_model.py:547-652).  With the fused engine the batches are
        prepared ahead on the prefetch thread and every batch is one C call.  ``n_rows`` (rows of every requested
        output over the whole pass, known to ``predict``): each batch's outputs are copied asynchronously into their
        place in ONE page-locked result array (the reference copies every output of every batch synchronously,
        torch_model.py:606); without it the outputs stay on the device and come back in one copy at the end."""
        self.model.eval()
        results = None
        with torch.no_grad():
            if self._engine is not None:
                cls = self.mode == "classification"
                chunks = None
                host, offs, sink = None, None, None
                for inputs, _, _ in _Prefetcher(self, generator, 2):
                    topo, n = inputs[1]._dcgc_topology, int(inputs[3])
                    out, probs, fp = self._engine.forward(topo, inputs[0], n, training=False, want_probs=cls)
                    if cls:
                        outs = [probs.view(n, self.n_tasks, self.n_classes), out.view(n, self.n_tasks, self.n_classes), fp]
                    else:
                        outs = [out, fp]
                    vals = [outs[i] for i in output_idx]
                    if chunks is None:
                        chunks = [[] for _ in vals]
                        if n_rows is not None:
                            nbytes = sum(r * v[0].numel() * 4 for r, v in zip(n_rows, vals))
                            if nbytes <= _PINNED_DIRECT_LIMIT:
                                host = [torch.empty((r,) + tuple(v.shape[1:]), dtype=torch.float32, pin_memory=True)
                                        for r, v in zip(n_rows, vals)]
                                offs = [0] * len(vals)
                            else:
                                sink = _StagedResult(self, [(r,) + tuple(v.shape[1:]) for r, v in zip(n_rows, vals)])
                    if sink is not None:
                        sink.add(vals)
                    elif host is not None:
                        for i, v in enumerate(vals):
                            if offs[i] + v.shape[0] > host[i].shape[0]:
                                raise ValueError("predict: more output rows than the dataset announced")
                            host[i][offs[i]:offs[i] + v.shape[0]].copy_(v, non_blocking=True)
                            offs[i] += v.shape[0]
                    else:
                        for c, v in zip(chunks, vals):
                            c.append(v)
                if chunks is None:
                    return []
                if sink is not None:
                    final = sink.finish()
                elif host is not None:
                    torch.cuda.current_stream().synchronize()
                    final = [h[:o].numpy() for h, o in zip(host, offs)]      # views keep the pinned tensors alive
                else:
                    final = [torch.cat(c, dim=0).cpu().numpy() for c in chunks]
                _check_f16_range(self)
                return final[0] if len(final) == 1 else final
            for batch in generator:
                inputs, _, _ = self._prepare_batch(batch)
                outs = self.model(inputs)
                vals = [outs[i].detach().cpu().numpy() for i in output_idx]
                if results is None:
                    results = [[] for _ in vals]
                for r, v in zip(results, vals):
                    r.append(v)
        if results is None:
            return []
        final = [np.concatenate(r, axis=0) for r in results]
        return final[0] if len(final) == 1 else final

    def predict(self, dataset, transformers=[], shard=None):
        """Predictions for the dataset (torch_model.py:731-761).  ``shard=(rank, world)`` restricts the pass to
        this rank's contiguous range of molecules (parallel.shard_range): sharded inference needs no
        communication, the caller concatenates the per-rank results in rank order."""
        if shard is not None:
            from .parallel import shard_range
            lo, hi = shard_range(len(dataset), int(shard[0]), int(shard[1]))
            dataset = dataset.select_range(lo, hi)
        gen = self.default_generator(dataset, mode='predict', deterministic=True, pad_batches=False)
        out = self._predict(gen, self._prediction_outputs, n_rows=[len(dataset)] * len(self._prediction_outputs))
        return _undo_transforms(out, transformers)

    def predict_on_generator(self, generator, transformers=[], output_types=None):
        """Predictions for batches from a generator (torch_model.py:654-690).  ``output_types``: names from
        ``self.output_types`` ('prediction', 'loss', 'variance', 'embedding') to return instead of the predictions."""
        idx = self._prediction_outputs
        if output_types is not None:
            wanted = [output_types] if isinstance(output_types, str) else list(output_types)
            idx = [i for i, t in enumerate(self.output_types) if t in wanted]
            if not idx:
                raise ValueError('This model cannot compute other outputs since no other output_types were specified.')
        return _undo_transforms(self._predict(generator, idx), transformers)

    def predict_on_batch(self, X, transformers=[]):
        ds = PackedDataset(X) if isinstance(X, PackedMols) else NumpyDataset(X)
        return self.predict(ds, transformers)

    def predict_uncertainty(self, dataset, masks=50):
        """(prediction, standard deviation) per sample and task (torch_model.py:763-817): ``masks`` passes are
        averaged, std = sqrt(E[y^2] - E[y]^2 + E[var]).  As in the torch reference the passes run the module in eval
        mode without its ``training`` argument (torch_model.py:597-603), so no dropout mask is drawn, all
        passes agree and the deviation is the aleatoric part sqrt(exp(log_var)); one pass is computed and reused."""
        if not self._variance_outputs:
            raise ValueError('This model cannot compute uncertainties')
        if len(self._variance_outputs) != len(self._prediction_outputs):
            raise ValueError('The number of variances must exactly match the number of outputs')
        if masks < 1:
            raise ValueError('masks must be positive')
        gen = self.default_generator(dataset, mode='uncertainty', deterministic=True, pad_batches=False)
        res = self._predict(gen, self._prediction_outputs + self._variance_outputs)
        k = len(self._prediction_outputs)
        pairs = [(p, np.sqrt(np.maximum(v, 0.0))) for p, v in zip(res[:k], res[k:])]     # E[y^2] - E[y]^2 = 0 here
        return pairs[0] if len(pairs) == 1 else pairs

    def predict_uncertainty_on_batch(self, X, masks=50):
        ds = PackedDataset(X) if isinstance(X, PackedMols) else NumpyDataset(X)
        return self.predict_uncertainty(ds, masks)

    def predict_embedding(self, dataset):
        """Untrimmed fingerprints, batch_size rows per batch, as the reference returns them."""
        gen = self.default_generator(dataset, mode='predict', deterministic=True, pad_batches=False)
        return self._predict(gen, self._embedding_outputs)

    def evaluate(self, dataset, metrics, transformers=[], per_task_metrics=False, use_sample_weights=False,
                 n_classes=2):
        """Model.evaluate (models.py:191-236) through the Evaluator's logic (``evaluate_model``)."""
        return evaluate_model(self, dataset, metrics, transformers, per_task_metrics, use_sample_weights, n_classes)

    # ------------------------------------------------------------------ checkpoints
    def get_checkpoints(self, model_dir=None):
        model_dir = model_dir or self.model_dir
        if not model_dir or not os.path.isdir(model_dir):
            return []
        files = [f for f in os.listdir(model_dir) if f.startswith("checkpoint") and f.endswith(".pt")]
        files.sort(key=lambda f: int(f[len("checkpoint"):-3]))
        return [os.path.join(model_dir, f) for f in files]

    def save_checkpoint(self, max_checkpoints_to_keep=5, model_dir=None):
        """checkpoint1.pt is the newest; older ones are shifted up (torch_model.py:996-1042)."""
        model_dir = model_dir or self.model_dir
        if model_dir is None:
            raise ValueError("model_dir is not set")
        if self._dp:
            from .parallel import rank
            if rank() != 0:                      # replicas are identical: one writer (ranks raced on the same files)
                return
        os.makedirs(model_dir, exist_ok=True)
        opt = self._engine.state_dict() if self._engine is not None else self._pytorch_optimizer.state_dict()
        data = {'model_state_dict': self.model.state_dict(), 'optimizer_state_dict': opt,
                'global_step': self._global_step}
        tmp = os.path.join(model_dir, 'temp_checkpoint.pt')
        torch.save(data, tmp)
        paths = [os.path.join(model_dir, 'checkpoint%d.pt' % (i + 1)) for i in range(max_checkpoints_to_keep)]
        if os.path.exists(paths[-1]):
            os.remove(paths[-1])
        for i in reversed(range(max_checkpoints_to_keep - 1)):
            if os.path.exists(paths[i]):
                os.rename(paths[i], paths[i + 1])
        os.rename(tmp, paths[0])

    def restore(self, checkpoint=None, model_dir=None):
        if checkpoint is None:
            cps = self.get_checkpoints(model_dir)
            if not cps:
                raise ValueError('No checkpoint found')
            checkpoint = cps[0]
        data = torch.load(checkpoint, map_location=self.device)
        self._bn_steps_pending = 0                      # (the loaded counters replace whatever was pending)
        self.model.load_state_dict(data['model_state_dict'])
        opt = data['optimizer_state_dict']
        if self._engine is not None:
            self._engine.load_state_dict(opt)           # torch.optim.Adam layout (or a round-1 slab checkpoint)
        elif 'state' in opt:
            self._pytorch_optimizer.load_state_dict(opt)
        else:
            raise ValueError("this checkpoint holds the raw moment slabs of the fused engine (written before the "
                             "optimizer state was saved in torch.optim.Adam's layout); restore it into a model "
                             "that uses the engine")
        self._global_step = data['global_step']

    def get_global_step(self):
        return self._global_step
