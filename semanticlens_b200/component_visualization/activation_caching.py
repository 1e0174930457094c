"""Collect, aggregate and cache top-k activations — B200 implementation.

Same classes, constructor arguments, attributes and on-disk format as the reference
(semanticlens/component_visualization/activation_caching.py): ``ActMax`` (:64-216), ``ActCache`` (:219-315),
``ActMaxCache`` (:318-534). What changed is where the work happens:

* the per-layer state ``(n_latents, n_collect)`` bf16 values + int64 ids lives in HBM for the whole sweep;
* the forward hook enqueues K1 (aggregate) + K2 (top-k merge) on the model's CUDA stream and returns — no
  ``clone()``, no ``.cpu()`` per layer and batch (reference: aggregators.py:61 + activation_caching.py:140);
* ``ActMax.activations`` / ``ActMax.sample_ids`` are CPU tensors when read (like the reference), mirrored lazily
  from the device state.

Tie order is canonical — (value desc, sample id asc, placeholders last) — where the reference's is whatever
``std::nth_element`` leaves; see DESIGN.md "tie-aware parity contract".
"""

from __future__ import annotations

import inspect
import logging
from collections import Counter, OrderedDict
from collections.abc import Callable
from contextlib import contextmanager
from pathlib import Path
from typing import Any

import safetensors
import safetensors.torch
import torch

from .. import _native, ops
from . import aggregators

logger = logging.getLogger(__name__)


DEFAULT_AGGREGATION_FUNCTION_MAP = {name: func for name, func in inspect.getmembers(aggregators, inspect.isfunction)}


class ActMax:
    """Streaming top-k of maximal activations for one layer (reference: activation_caching.py:64-216).

    Parameters
    ----------
    n_collect : int
        Number of top activations kept per latent (k). ``0`` is legal and yields ``(n_latents, 0)`` tensors.
    n_latents : int, optional
        Number of latents; inferred from the first batch when omitted.
    """

    def __init__(self, n_collect: int, n_latents: int | None = None):
        self.n_collect = n_collect
        self.n_latents = n_latents
        self.is_setup = False
        self._dev_vals: torch.Tensor | None = None  # (C, k) bf16 on the GPU while sweeping
        self._dev_ids: torch.Tensor | None = None
        self._cpu_vals: torch.Tensor | None = None  # lazily mirrored copy handed to readers
        self._cpu_ids: torch.Tensor | None = None
        self._scratch: torch.Tensor | None = None

        if n_latents is not None:
            self._setup_tensors()

    # -- state ----------------------------------------------------------------------------------
    def _setup_tensors(self):
        """Fresh state: values -0.0 (bf16), ids -1 (reference :101-110)."""
        self._cpu_vals = -torch.zeros(self.n_latents, self.n_collect, dtype=torch.bfloat16)
        self._cpu_ids = -torch.ones(self.n_latents, self.n_collect, dtype=torch.int64)
        self._dev_vals = self._dev_ids = None
        self.is_setup = True

    def _device_state(self, device) -> tuple[torch.Tensor, torch.Tensor]:
        if self._dev_vals is None or self._dev_vals.device != torch.device(device):
            src_v = self._cpu_vals if self._dev_vals is None else self._dev_vals
            src_i = self._cpu_ids if self._dev_ids is None else self._dev_ids
            self._dev_vals = src_v.to(device).contiguous()
            self._dev_ids = src_i.to(device).contiguous()
        self._cpu_vals = self._cpu_ids = None  # the device copy is now the truth
        return self._dev_vals, self._dev_ids

    @property
    def activations(self) -> torch.Tensor:
        """(n_latents, n_collect) bf16, sorted descending, on the CPU."""
        if self._cpu_vals is None:
            self._cpu_vals = self._dev_vals.cpu()
        return self._cpu_vals

    @activations.setter
    def activations(self, value: torch.Tensor):
        self._cpu_vals = value.detach().to("cpu", torch.bfloat16).contiguous()
        if self._cpu_ids is None and self._dev_ids is not None:
            self._cpu_ids = self._dev_ids.cpu()
        self._dev_vals = self._dev_ids = None

    @property
    def sample_ids(self) -> torch.Tensor:
        """(n_latents, n_collect) int64 dataset indices (-1 = empty slot), on the CPU."""
        if self._cpu_ids is None:
            self._cpu_ids = self._dev_ids.cpu()
        return self._cpu_ids

    @sample_ids.setter
    def sample_ids(self, value: torch.Tensor):
        self._cpu_ids = value.detach().to("cpu", torch.int64).contiguous()
        if self._cpu_vals is None and self._dev_vals is not None:
            self._cpu_vals = self._dev_vals.cpu()
        self._dev_vals = self._dev_ids = None

    def device_tensors(self, device=None) -> tuple[torch.Tensor, torch.Tensor]:
        """The device-resident state (values bf16, ids int64) — no host copy."""
        if device is None:
            device = self._dev_vals.device if self._dev_vals is not None else torch.device("cuda")
        return self._device_state(device)

    def finalize(self):
        """Mirror the device state to the host (one small D2H per layer, at the end of the sweep)."""
        if self.is_setup:
            _ = self.activations, self.sample_ids

    # -- updates --------------------------------------------------------------------------------
    def update(self, acts: torch.Tensor, sample_ids: torch.Tensor):
        """Merge a (batch, n_latents) aggregate with its sample ids into the top-k (reference :112-141).

        ``acts`` may live on the CPU (as the reference's aggregators return it); it is copied to the GPU —
        the selection itself always runs in the K2 kernel.
        """
        assert acts.ndim == 2
        _native.load(require_device=True)  # fail loudly: there is no CPU top-k in this package
        if not self.is_setup:
            self.n_latents = acts.shape[1]
            self._setup_tensors()
        device = acts.device if acts.is_cuda else torch.device("cuda")
        vals, ids = self._device_state(device)
        ops.topk_update(acts.detach().to(device), vals, ids, ids=sample_ids)

    def update_from_map(self, outs: torch.Tensor, op: int, kind: str, token: int, id_base: int):
        """Fused hook path: K1 aggregate + K2 merge straight from the hooked map, ids = id_base + arange(B)."""
        n_latents = outs.shape[1] if kind == "conv" else outs.shape[2]
        if not self.is_setup:
            self.n_latents = n_latents
            self._setup_tensors()
        assert n_latents == self.n_latents, (n_latents, self.n_latents)
        vals, ids = self._device_state(outs.device)
        self._scratch = ops.agg_topk_update(outs, op, kind, token, id_base, vals, ids, self._scratch)

    @property
    def alive_latents(self) -> torch.Tensor:
        """Indices of latents with any non-zero activation (reference :143-156)."""
        if not self.is_setup:
            return torch.tensor([], dtype=torch.int64)
        return torch.where(self.activations.abs().sum(dim=1) > 0)[0]

    # -- persistence ---------------------------------------------------------------------------
    # On-disk contract shared with the reference (activation_caching.py:158-216) so caches interchange: one safetensors
    # file with tensors "activations" (bf16) and "sample_ids" (int64) and string metadata carrying at least
    # n_collect and n_latents. Everything else about how the file is produced / parsed is this module's own.
    def store(self, file_path: str | Path, metadata: dict[str, str] | None = None):
        if not self.is_setup:
            logger.warning("ActMax.store: nothing collected yet, %s not written", file_path)
            return
        _write_state_file(Path(file_path), self.activations, self.sample_ids, metadata)

    @classmethod
    def load(cls, file_path: str | Path) -> ActMax:
        meta, vals, ids = _read_state_file(Path(file_path))
        if meta is None or "n_collect" not in meta or "n_latents" not in meta:
            raise ValueError(f"{file_path} carries no n_collect / n_latents metadata: not an ActMax file")
        state = cls(n_collect=int(meta["n_collect"]), n_latents=int(meta["n_latents"]))
        state._cpu_vals = vals.to(torch.bfloat16).contiguous()
        state._cpu_ids = ids.to(torch.int64).contiguous()
        state._dev_vals = state._dev_ids = None
        return state


def _write_state_file(path: Path, vals: torch.Tensor, ids: torch.Tensor, metadata: dict[str, str] | None) -> None:
    """Write to a sibling temporary name and rename, so a reader (another rank, a resumed run) never sees half a file."""
    tmp = path.with_name(path.name + ".partial")
    safetensors.torch.save_file({"activations": vals, "sample_ids": ids}, str(tmp), metadata=metadata)
    tmp.replace(path)
    logger.debug("wrote %s", path)


def _read_state_file(path: Path, header_only: bool = False):
    with safetensors.safe_open(str(path), framework="pt") as f:
        meta = f.metadata()
        if header_only:
            return meta, None, None
        return meta, f.get_tensor("activations"), f.get_tensor("sample_ids")


class ActCache:
    """Forward-hook plumbing (reference: activation_caching.py:219-315)."""

    def __init__(self, layer_names: list[str]):
        self.layer_names = layer_names
        self.cache: dict[str, Any] = OrderedDict()
        self.handles: list[torch.utils.hooks.RemovableHandle] = []

    def _get_hook(self, name: str) -> Callable:
        def hook_fn(module, ins, outs):
            self.cache[name] = outs.detach().cpu()

        return hook_fn

    def _register_hooks(self, model: torch.nn.Module):
        for name, module in model.named_modules():
            if name in self.layer_names:
                self.handles.append(module.register_forward_hook(self._get_hook(name)))

    def _finalize(self):
        pass

    @contextmanager
    def hook_context(self, model: torch.nn.Module):
        """Register hooks, yield, always remove them (exception safe) and finalise."""
        self._register_hooks(model)
        try:
            yield
        finally:
            for handle in self.handles:
                handle.remove()
            self.handles.clear()
            self._finalize()


class ActMaxCache(ActCache):
    """Per-layer aggregation + streaming top-k behind forward hooks (reference: activation_caching.py:318-534)."""

    def __init__(self, layer_names: list[str], aggregation_fn: Callable, n_collect: int):
        super().__init__(layer_names)
        self.aggregation_fn = aggregation_fn
        self.n_collect = n_collect
        self.sample_idx_counter = Counter()

        agg_fn_name = getattr(self.aggregation_fn, "__name__", None)
        if agg_fn_name is None or agg_fn_name == "<lambda>":
            raise ValueError("Aggregation function must be a defined function, not a lambda.")
        self.agg_fn_name = agg_fn_name

        self.cache: dict[str, ActMax] = {name: ActMax(n_collect=n_collect) for name in layer_names}

    def __getitem__(self, layer_name: str) -> ActMax:
        return self.cache[layer_name]

    def __iter__(self):
        return iter(self.cache.values())

    def _get_hook(self, layer_name: str) -> Callable:
        """Hook = K1 + K2 on the model's stream, ids numbered in iteration order (reference :388-418)."""
        fn = self.aggregation_fn
        fused = hasattr(fn, "_slb_op")

        def hook_fn(module, ins, outs):
            if fused and isinstance(outs, torch.Tensor) and outs.is_cuda:
                want = 4 if fn._slb_kind == "conv" else 3
                if outs.ndim != want:
                    fn(outs)  # raises the reference's ValueError for a wrong rank
                if fn._slb_op == _native.AGG_TOKEN and not -outs.shape[1] <= fn._slb_token < outs.shape[1]:
                    fn(outs)  # raises the IndexError python indexing gives the reference
                batch_size = outs.shape[0]
                self.cache[layer_name].update_from_map(
                    outs, fn._slb_op, fn._slb_kind, fn._slb_token, self.sample_idx_counter[layer_name]
                )
                self.sample_idx_counter[layer_name] += batch_size
                return

            # user-defined aggregation function (or a CPU model): same contract as the reference
            aggregated_acts = fn(outs)
            batch_size = aggregated_acts.shape[0]

            assert aggregated_acts.ndim == 2, "Something is wrong with the aggregation_fn"

            sample_ids = torch.arange(
                self.sample_idx_counter[layer_name], self.sample_idx_counter[layer_name] + batch_size
            )
            self.sample_idx_counter[layer_name] += batch_size
            self.cache[layer_name].update(aggregated_acts, sample_ids)

        return hook_fn

    def _finalize(self):
        """End of the hook context: bring the (tiny) per-layer states to the host once."""
        for act_max in self.cache.values():
            act_max.finalize()

    def __repr__(self) -> str:
        agg_name = getattr(self.aggregation_fn, "__name__", "custom_function")
        return f"ActMaxCache(layers={list(self.layer_names)}, aggregation_fn='{agg_name}', n_collect={self.n_collect})"

    @property
    def metadata(self) -> dict[str, str]:
        """Key order matters: Lens joins the values into the concept-DB file name (reference lens.py:308-316)."""
        return {
            "aggregation_fn_name": self.agg_fn_name,
            "n_collect": str(self.n_collect),
            "layer_names": str(list(self.cache)),
        }

    # File grammar shared with the reference (:434-465, :495-503): "<agg fn>-<n_collect>-<layer>.safetensors", metadata
    # keys aggregation_fn_name / n_collect / n_latents / layer_name.
    def _layer_path(self, directory: Path, layer_name: str) -> Path:
        return directory / f"{self.agg_fn_name}-{self.n_collect}-{layer_name}.safetensors"

    def store(self, directory: Path | str):
        directory = Path(directory)
        directory.mkdir(parents=True, exist_ok=True)
        written = 0
        for layer_name, state in self.cache.items():
            if not state.is_setup:
                logger.warning("layer '%s' saw no batch; no file written for it", layer_name)
                continue
            state.store(self._layer_path(directory, layer_name), metadata={
                "aggregation_fn_name": self.agg_fn_name,
                "n_collect": str(self.n_collect),
                "n_latents": str(state.n_latents),
                "layer_name": layer_name,
            })
            written += 1
        logger.info("act-max cache: %d layer file(s) written to %s", written, directory)

    def load(self, directory: Path | str):
        """All layers or nothing: every file is located and its header validated before any state is replaced. A missing
        directory / file or a header written for another aggregation function or n_collect is reported as
        ``FileNotFoundError`` — the signal ``run()`` and the constructor treat as "cache miss" (reference :467-534)."""
        directory = Path(directory)
        if not directory.is_dir():
            raise FileNotFoundError(f"no act-max cache at {directory}")
        paths = {}
        for layer_name in self.layer_names:
            path = self._layer_path(directory, layer_name)
            if not path.is_file():
                raise FileNotFoundError(f"act-max cache has no file for layer '{layer_name}': {path}")
            meta = _read_state_file(path, header_only=True)[0] or {}
            problems = []
            if meta.get("aggregation_fn_name") != self.agg_fn_name:
                problems.append(f"aggregation_fn_name={meta.get('aggregation_fn_name')!r}, wanted {self.agg_fn_name!r}")
            if str(meta.get("n_collect")) != str(self.n_collect):
                problems.append(f"n_collect={meta.get('n_collect')!r}, wanted {self.n_collect}")
            if problems:
                logger.warning("ignoring %s: %s", path, "; ".join(problems))
                raise FileNotFoundError(f"act-max cache file does not match this configuration: {path}")
            paths[layer_name] = path
        loaded = {layer_name: ActMax.load(path) for layer_name, path in paths.items()}
        self.cache.update(loaded)
        if loaded:
            logger.info("act-max cache: %d layer(s) loaded from %s", len(loaded), directory)
        else:
            logger.warning("act-max cache: no layers requested, nothing loaded from %s", directory)
