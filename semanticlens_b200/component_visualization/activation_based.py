"""Activation-based component visualization — B200 implementation.

Drop-in for ``semanticlens.component_visualization.activation_based.ActivationComponentVisualizer``
(reference: activation_based.py:41-560): same constructor, ``run`` / ``_run`` / ``_compute_concept_db`` /
``_embed_vision_dataset`` / ``get_max_reference``, same cache directory grammar (the class name is part of the
path, :295), same exceptions and warnings.

Differences that are not API-visible:

* the sweep never synchronises the host: hooks enqueue K1+K2 on the model's stream (activation_caching.py), input
  batches are copied host->device from pinned memory with ``non_blocking=True`` and the logits are not copied back
  (reference :352 does ``model(images.to(device)).cpu()``);
* when ``torch.distributed`` is initialised with world_size R > 1, rank r sweeps and embeds the contiguous index
  balanced shard of the index range; the per-rank top-k states are exchanged with ONE all-gather and merged by the K2
  list-merge kernel, then one more all-gather moves only the embedding rows the merged top-k refers to
  (``semanticlens_b200.distributed``);
* embeddings stay in HBM and ``embeds[sample_ids]`` is the K5 gather kernel; the concept DB is returned on the CPU
  like the reference unless ``output_device`` is set.
"""

from __future__ import annotations

import logging
import warnings
from pathlib import Path

import torch
from torch import nn
from tqdm import tqdm

from .. import distributed as sdist
from .. import ops
from ..utils.helper import get_fallback_name
from . import aggregators
from .activation_caching import ActMaxCache
from .base import AbstractComponentVisualizer

logger = logging.getLogger(__name__)


class _SlicedBatches:
    """Batches of a dataset that can hand out a whole index range at once (``dataset.get_batch(lo, hi)``), e.g. one
    backed by a single pinned host tensor: no per-item ``__getitem__`` + collate copy on the host."""

    def __init__(self, dataset, lo, hi, batch_size):
        self.dataset, self.lo, self.hi, self.batch_size = dataset, lo, hi, batch_size

    def __len__(self):
        return -(-(self.hi - self.lo) // self.batch_size) if self.hi > self.lo else 0

    def __iter__(self):
        for a in range(self.lo, self.hi, self.batch_size):
            yield self.dataset.get_batch(a, min(a + self.batch_size, self.hi))


def _batches(dataset, shard, batch_size, num_workers, device, collate_fn=None):
    """Iteration order of the reference (DataLoader, shuffle=False, :344-349 / :414-422) over this rank's shard."""
    if hasattr(dataset, "get_batch"):
        return _SlicedBatches(dataset, shard.lo, shard.hi, batch_size)
    if shard.world > 1:
        dataset = torch.utils.data.Subset(dataset, range(shard.lo, shard.hi))
    kw = dict(collate_fn=collate_fn) if collate_fn is not None else dict(pin_memory=device.type == "cuda")
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=num_workers, **kw)


class _DevicePrefetcher:
    """Iterate ``batches`` with the host->device copy of batch i+1 in flight (on a copy stream) while batch i is being
    computed, so the H2D time disappears behind the probed-model forward / the tower.

    Pinned host tensors are copied into one of TWO device buffers owned by the prefetcher (allocated once: no caching-
    allocator traffic, no ``cudaMalloc`` in the loop); events order the hand-over in both directions (copy done ->
    consumer may read; consumer done -> buffer may be overwritten). Unpinned tensors and non-tensor items (lists of PIL
    images) pass through untouched. Order and contents are exactly those of ``batches``."""

    def __init__(self, batches, device):
        self.batches, self.device = batches, torch.device(device)

    def __len__(self):
        return len(self.batches)

    @staticmethod
    def _tensor_of(item):
        if isinstance(item, torch.Tensor):
            return item
        if isinstance(item, (tuple, list)) and len(item) == 2 and isinstance(item[0], torch.Tensor) and item[0].ndim >= 3:
            return item[0]  # an (images, labels) batch of dataset_model
        return None

    def __iter__(self):
        if self.device.type != "cuda":
            yield from self.batches
            return
        dev = self.device
        copy_stream = torch.cuda.Stream(dev)
        bufs: list = [None, None]
        reusable: list = [None, None]  # event on the compute stream after which buffer k may be overwritten

        def stage(item, k):
            t = self._tensor_of(item)
            if t is None or t.is_cuda or not t.is_pinned():
                return item, None
            cur_stream = torch.cuda.current_stream(dev)
            if bufs[k] is None or bufs[k].numel() < t.numel() or bufs[k].dtype != t.dtype:
                bufs[k] = torch.empty(t.numel(), dtype=t.dtype, device=dev)
                fresh = torch.cuda.Event()
                fresh.record(cur_stream)  # the block may have been used by earlier work of the compute stream
                copy_stream.wait_event(fresh)
            if reusable[k] is not None:
                copy_stream.wait_event(reusable[k])
            dst = bufs[k][: t.numel()].view(t.shape)
            with torch.cuda.stream(copy_stream):
                dst.copy_(t, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            out = dst if isinstance(item, torch.Tensor) else type(item)((dst, *item[1:]))
            return out, ready

        it = iter(self.batches)
        try:
            nxt = stage(next(it), 0)
        except StopIteration:
            return
        k = 0
        while nxt is not None:
            cur, ready = nxt
            k_cur, k = k, k ^ 1
            try:
                nxt = stage(next(it), k)  # enqueue the next copy before handing out the current batch
            except StopIteration:
                nxt = None
            if ready is not None:
                torch.cuda.current_stream(dev).wait_event(ready)
            yield cur
            if ready is not None:  # the consumer asked for the next batch: everything it enqueued on `cur` is in the stream
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(dev))
                reusable[k_cur] = done


class MissingNameWarning(UserWarning):
    """A model or dataset has no ``.name``; a fallback derived from its ``repr`` names the cache directory."""


class ActivationComponentVisualizer(AbstractComponentVisualizer):
    """Find, per component of the chosen layers, the ``num_samples`` dataset items that activate it most.

    Parameters mirror the reference (activation_based.py:124-134).
    """

    AGGREGATION_DEFAULTS = {
        "max": aggregators.aggregate_conv_max,
        "mean": aggregators.aggregate_conv_mean,
    }

    def __init__(
        self,
        model: nn.Module,
        dataset_model,
        dataset_fm,
        layer_names: list[str],
        num_samples: int,
        device=None,
        aggregate_fn=None,
        cache_dir: str | None = None,
        accelerate: bool | None = None,
    ):
        self.model = model
        # opt-in: run a torchvision-style ResNet on the package's own convolution kernels instead of torch's (probed.py).
        # None = the SLB_ACCEL_FORWARD environment variable. Not part of the cache key: the maps agree to ~1e-5.
        self.accelerate = accelerate
        self._accel_forward = None
        self.dataset = dataset_model
        self.dataset_fm = dataset_fm
        self.output_device = "cpu"  # where _compute_concept_db leaves the (C, k, D) tensors
        self.exchange = "winners"  # distributed embed exchange: "winners" (rows the top-k refers to) or "all"
        self.show_progress = True
        self._init_cache_dir(cache_dir)
        self._validate_args()

        self.layer_names = layer_names
        self._check_layers()

        device = device or next(model.parameters()).device
        self.model.to(device)

        if aggregate_fn is None:
            logger.warning(f"No aggregation_fn provided using default: {aggregators.aggregate_conv_mean.__name__}")
            aggregate_fn = aggregators.aggregate_conv_mean

        self.actmax_cache = ActMaxCache(self.layer_names, n_collect=num_samples, aggregation_fn=aggregate_fn)

        if self.caching:
            try:
                self.actmax_cache.load(self.storage_dir)
                logger.info(f"Results loaded from {self.storage_dir}")
            except FileNotFoundError:
                logger.info(f"Results will be stored in {self.storage_dir}")

    # -- argument checks (reference :187-229) ------------------------------------------------------
    def _validate_args(self):
        for obj, what in ((self.model, "Model"), (self.dataset, "Dataset")):
            if hasattr(obj, "name"):
                continue
            fallback = get_fallback_name(obj)
            if self.caching:
                warnings.warn(
                    f"{what} does not have a name attribute, which is required for reliable caching.\n"
                    f"Using a fallback name: {fallback}.",
                    MissingNameWarning,
                    stacklevel=3,
                )
            obj.name = fallback

        if len(self.dataset) != len(self.dataset_fm):
            raise ValueError(
                "Model and foundation model datasets should have the same length.",
                (len(self.dataset), len(self.dataset_fm)),
            )

    def _check_layers(self):
        modules = dict(self.model.named_modules())
        for layer in self.layer_names:
            if layer not in modules:
                raise ValueError(f"Layer '{layer}' not found in model.")

    def _check_layer_name(self, layer_name):
        if layer_name not in self.layer_names:
            raise ValueError(f"Layer '{layer_name}' not found in model layers: {self.layer_names}")

    def _init_cache_dir(self, cache_dir):
        if cache_dir is None:
            logger.warning("No cache dir provided. Results will not be cached!")
            self._cache_root = None
        else:
            self._cache_root = Path(cache_dir)
            self._cache_root.mkdir(parents=True, exist_ok=True)

    # -- plugin properties ---------------------------------------------------------------------------
    @property
    def device(self):
        return next(self.model.parameters()).device

    def to(self, device):
        return self.model.to(device)

    @property
    def caching(self) -> bool:
        return self._cache_root is not None

    @property
    def storage_dir(self):
        assert self._cache_root, "No cache dir provided"
        return self._cache_root / self.__class__.__name__ / self.dataset.name / self.model.name

    @property
    def metadata(self) -> dict[str, str]:
        return {**self.actmax_cache.metadata, "dataset": self.dataset.name, "model": self.model.name}

    # -- collect --------------------------------------------------------------------------------------
    def run(self, batch_size=32, num_workers=0):
        """Sweep ``dataset_model`` (or load the cached result) -> ``{layer: ActMax}`` (reference :309-339).

        Distributed: whether the cache is warm is decided by rank 0 alone and its states are sent to the other ranks, so
        a cache directory only rank 0 can see cannot make the ranks disagree about entering the sweep's collectives."""
        if not self.caching:
            logger.debug("No cache root provided, running computation...")
            return self._run(batch_size=batch_size, num_workers=num_workers)
        rank, world_size = sdist.world()
        hit = False
        if rank == 0:
            try:
                self.actmax_cache.load(self.storage_dir)
                hit = True
            except FileNotFoundError:
                logger.debug(f"Activation maximization cache not found at {self.storage_dir}. Running computation...")
        if world_size > 1:
            hit = sdist.agree(hit, self.device)
            if hit:
                sdist.share_actmax_from_rank0(self.actmax_cache)
        if hit:
            return self.actmax_cache.cache
        return self._run(batch_size=batch_size, num_workers=num_workers)

    def _probed_forward(self, device):
        """``self.model`` itself (the reference's ``self.model(x)``), or — opt-in, torchvision-style ResNets only — the
        B200 forward that produces the hooked maps with the package's convolution kernels (probed.AcceleratedResNet)."""
        from .. import probed

        if not probed.accel_requested(self.accelerate):
            return self.model
        if self._accel_forward is None or self._accel_forward.device != torch.device(device):
            self._accel_forward = probed.accelerated_forward(self.model, device)
        return self._accel_forward

    @torch.no_grad()
    def _run(self, batch_size: int = 64, num_workers: int = 0):
        """The activation sweep (reference :341-358), image-sharded across ranks when distributed."""
        sdist.require_items_per_rank(len(self.dataset))
        shard = sdist.image_shard(len(self.dataset))
        device = self.device
        dataloader = _batches(self.dataset, shard, batch_size, num_workers, device)
        if shard.world > 1:
            # ids are positions in iteration order; a shard starts at its offset, on a fresh state
            for layer in self.layer_names:
                self.actmax_cache.cache[layer] = type(self.actmax_cache.cache[layer])(self.actmax_cache.n_collect)
                self.actmax_cache.sample_idx_counter[layer] = shard.lo
        forward = self._probed_forward(device)
        with self.actmax_cache.hook_context(self.model):
            for images, _ in tqdm(
                _DevicePrefetcher(dataloader, device), total=len(dataloader), desc="Collecting ActMax",
                disable=not self.show_progress,
            ):
                forward(images.to(device, non_blocking=True))  # hooks enqueue K1+K2; nothing is copied back

        if shard.world > 1:
            sdist.merge_actmax_across_ranks(self.actmax_cache, device)

        if self._cache_root and shard.rank == 0:
            self.actmax_cache.store(self.storage_dir)
            logger.debug(f"Stored activation maximization cache at {self.storage_dir}")

        return self.actmax_cache.cache

    # -- embed + gather -------------------------------------------------------------------------------
    @torch.no_grad()
    def _compute_concept_db(self, fm, batch_size=32, **kwargs):
        """``{layer: embeds[sample_ids]}`` with embeds = fm image embeddings of ``dataset_fm`` (reference :360-390).

        Distributed: each rank embeds its shard and only the rows the merged top-k refers to are exchanged
        (``sdist.exchange_winner_rows``); set ``exchange = "all"`` to all-gather the whole table instead."""
        self.run(batch_size=batch_size, **kwargs)
        rank, world_size = sdist.world()
        ids = {name: self.get_max_reference(name) for name in self.layer_names}

        if world_size > 1 and self.exchange == "winners":
            local, shard = self._embed_vision_dataset(fm, batch_size, _gather=False, **kwargs)
            dev_ids = [ids[name].to(local.device) for name in self.layer_names]
            embeds, remapped = sdist.exchange_winner_rows(local, shard, dev_ids)
            ids = dict(zip(self.layer_names, remapped))
        else:
            embeds = self._embed_vision_dataset(fm, batch_size, **kwargs)

        concept_db = dict()
        to_host = torch.device(self.output_device).type == "cpu"
        for layer_name in self.layer_names:
            if embeds.is_cuda:
                table = embeds if embeds.dtype == torch.float32 else embeds.float()
                db = ops.gather_rows(table, ids[layer_name])  # K5; python-negative semantics: id -1 -> last image
                if to_host:
                    # asynchronous D2H into pinned memory (the next layer's gather runs meanwhile); one sync below
                    host = torch.empty(db.shape, dtype=db.dtype, pin_memory=True)
                    host.copy_(db, non_blocking=True)
                    concept_db[layer_name] = host
                else:
                    concept_db[layer_name] = db.to(self.output_device)
            else:
                concept_db[layer_name] = embeds[ids[layer_name]]
        if embeds.is_cuda and to_host:
            torch.cuda.current_stream(embeds.device).synchronize()
        return concept_db

    def _embed_vision_dataset(self, fm, batch_size, _gather=True, **kwargs):
        """Embed every item of ``dataset_fm`` -> (N, D) fp32 (reference :392-433); stays on the GPU.

        ``_gather=False`` (distributed, internal) returns ``(this rank's rows, shard)`` without the all-gather."""
        fm.to(self.device)

        def item_list_collate(batch):
            # dataset_fm yields PIL images (or uint8 CHW tensors); fm.preprocess is applied per batch
            if isinstance(batch[0], (tuple, list)):
                return [item[0] for item in batch]
            return list(batch)

        sdist.require_items_per_rank(len(self.dataset_fm), "foundation-model dataset")
        shard = sdist.image_shard(len(self.dataset_fm))
        loader = _batches(self.dataset_fm, shard, batch_size, kwargs.get("num_workers", 0), self.device,
                          collate_fn=item_list_collate)
        embeds = []
        with tqdm(total=shard.hi - shard.lo, desc="Embedding Dataset", disable=not self.show_progress) as pbar:
            for items in _DevicePrefetcher(loader, self.device):
                inputs = fm.preprocess(items)
                embeds.append(fm.encode_image(inputs))
                pbar.update(len(items))
        if embeds:
            embeds = torch.cat(embeds)
        else:
            embeds = torch.empty((0, 0), dtype=torch.float32, device=self.device)
        assert embeds.shape[0] == shard.hi - shard.lo, "Number of embeddings does not match number of ids!"

        if not _gather:
            return embeds, shard
        if shard.world > 1:
            embeds = sdist.all_gather_rows(embeds, shard, len(self.dataset_fm))

        assert embeds.shape[0] == len(self.dataset_fm), "Number of embeddings does not match number of ids!"
        return embeds

    def get_max_reference(self, layer_name) -> torch.Tensor:
        """(n_components, n_samples) int64 dataset indices of the top activating samples (reference :435-451)."""
        self._check_layer_name(layer_name)
        return self.actmax_cache.cache[layer_name].sample_ids

    def visualize_components(self, *args, **kwargs):
        """Matplotlib grid plotting (reference :453-543) is outside the concept-DB build path."""
        raise NotImplementedError(
            "visualize_components is plotting-only and is not part of the B200 concept-database path; "
            "use get_max_reference(layer) and plot the dataset items yourself."
        )
