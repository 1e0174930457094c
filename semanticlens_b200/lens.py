"""``Lens`` — orchestrator of the concept-database path; drop-in for ``semanticlens.lens`` (reference lens.py:27-480).

Same functions, signatures and cache-file grammar as the reference. What changes underneath:
``cv._compute_concept_db(fm)`` is the B200 sweep + embed + gather, and ``_probe`` / ``eval_*`` call the libslb200 score
kernels (``semanticlens_b200.scores``).
"""

from __future__ import annotations

import logging

import torch
from safetensors.torch import load_file, save_file
from tqdm.auto import tqdm

from . import distributed as sdist
from .component_visualization.base import AbstractComponentVisualizer
from .foundation_models.base import AbstractVLM
from .scores import clarity_score, polysemanticity_score, redundancy_score, similarity_score
from .utils.helper import get_fallback_name

logger = logging.getLogger(__name__)


def compute_concept_db(cv: AbstractComponentVisualizer, fm: AbstractVLM):
    """Stateless form (reference lens.py:27-56): ``{layer: (n_components, n_samples, embed_dim)}``."""
    return cv._compute_concept_db(fm)


def text_probing(fm, query, aggregated_concept_db, templates=None, batch_size=None):
    """Cosine similarity of text-query embeddings with an aggregated concept DB (reference lens.py:59-121)."""
    queries = query if isinstance(query, list) else [query]
    query_embeds = _embed_text_probes(fm, queries, templates, batch_size)
    assert query_embeds.ndim == 2
    assert query_embeds.shape[0] == len(queries)
    return _probe(query_embeds, aggregated_concept_db)


def image_probing(fm, query, aggregated_concept_db):
    """Cosine similarity of image-query embeddings with an aggregated concept DB (reference lens.py:124-162)."""
    with torch.no_grad():
        query_embed = fm.encode_image(fm.preprocess(query).to(fm.device)).cpu()
    query_embed = query_embed.mean(0)[None] if query_embed.shape[0] > 1 else query_embed
    return _probe(query_embed, aggregated_concept_db)


@torch.no_grad()
def _embed_text_probes(fm, query: list[str], templates, batch_size):
    """Text-query embeddings (reference lens.py:166-203).

    Without templates: one embedding per query. With templates: every (template, query) prompt is embedded, the
    embedding of the same template formatted with "" is subtracted, and the result is averaged per query. The prompt
    list is built template-major (all queries under template 0, then template 1, ...) while the regrouping reads it as
    ``(query, template)`` blocks — the reference's behaviour (lens.py:174 vs :196-199), which mixes queries unless one
    of the two lists has a single entry. It is reproduced on purpose so results stay comparable."""

    def embed(prompts):
        return fm.encode_text(fm.tokenize(prompts).to(fm.device)).cpu()

    if not templates:
        return embed(query)
    prompts = [template.format(q) for template in templates for q in query]
    step = batch_size or len(prompts)
    starts = range(0, len(prompts), step)
    if len(starts) > 1:
        starts = tqdm(starts, desc="text embedding ...", leave=False)
    rows = torch.cat([embed(prompts[b0 : b0 + step]) for b0 in starts], dim=0)
    baseline = embed([template.format("") for template in templates])
    n_q, n_t = len(query), len(templates)
    return (rows.view(n_q, n_t, rows.shape[-1]) - baseline.view(1, n_t, baseline.shape[-1])).mean(dim=1)


@torch.no_grad()
def _probe(query: torch.Tensor, aggregated_concept_db):
    """K6 (``similarity_score``) for one aggregated concept DB or a dict of them (reference lens.py:207-214)."""

    def one(db):
        return similarity_score(query.to(db.device), db)

    if isinstance(aggregated_concept_db, dict):
        return {layer: one(db) for layer, db in aggregated_concept_db.items()}
    return one(aggregated_concept_db)


def _per_layer(score_fn, db):
    """The reference's eval_* dispatch (lens.py:391-480): a tensor is scored directly, a dict layer by layer."""
    if isinstance(db, dict):
        return {layer: score_fn(t) for layer, t in db.items()}
    return score_fn(db)


class Lens:
    """Holds the foundation model; computes / caches concept DBs and evaluates them (reference lens.py:217-480)."""

    def __init__(self, fm, device=None):
        self.fm = fm
        self.device = device if device is not None else fm.device
        fm.to(self.device)
        if not hasattr(fm, "name"):
            fm.name = get_fallback_name(fm)  # names the concept-DB cache directory
            logger.debug("foundation model has no .name; using %s", fm.name)

    def concept_db_path(self, cv: AbstractComponentVisualizer):
        """Cache file of ``cv``'s concept DB under this foundation model — the reference's grammar (lens.py:308-316):
        ``<storage_dir>/concept_database/<fm.name>/concept_db-<agg fn>-<n_collect>-<layer list>.safetensors``."""
        tags = [value for key, value in cv.metadata.items() if key not in ("dataset", "model")]
        return cv.storage_dir / "concept_database" / self.fm.name / ("concept_db-" + "-".join(tags) + ".safetensors")

    def compute_concept_db(self, cv: AbstractComponentVisualizer, **kwargs) -> dict[str, torch.Tensor]:
        """Concept DB of ``cv`` under ``self.fm``; with ``cv.caching`` it is read from / written to
        :meth:`concept_db_path`.

        Under ``torch.distributed`` rank 0 decides hit or miss for everyone and is the only writer, and a barrier
        follows the write, so no rank reads a half-written file or skips the collectives of
        ``cv._compute_concept_db``. The cache directory must be visible to every rank that wants to read it."""
        if not cv.caching:
            logger.debug("Caching is not enabled. Computing Concept DB")
            return cv._compute_concept_db(self.fm, **kwargs)
        rank, _ = sdist.world()
        path = self.concept_db_path(cv)
        if sdist.agree(path.exists() if rank == 0 else False, getattr(cv, "device", None)):
            logger.debug("Loading concept DB from cache %s", path)
            return load_file(filename=path)
        concept_db = cv._compute_concept_db(self.fm, **kwargs)
        if rank == 0:
            path.parent.mkdir(parents=True, exist_ok=True)
            save_file(tensors={layer: t.cpu().contiguous() for layer, t in concept_db.items()}, filename=path)
            logger.debug("Saved concept DB to cache %s", path)
        sdist.barrier()
        return concept_db

    def text_probing(self, query, aggregated_concept_db, templates=None, batch_size=None):
        return text_probing(self.fm, query, aggregated_concept_db, templates, batch_size)

    def image_probing(self, query, aggregated_concept_db):
        return image_probing(self.fm, query, aggregated_concept_db)

    def eval_clarity(self, concept_db):
        return _per_layer(clarity_score, concept_db)

    def eval_redundancy(self, aggregated_concept_db):
        return _per_layer(redundancy_score, aggregated_concept_db)

    def eval_polysemanticity(self, concept_db):
        return _per_layer(polysemanticity_score, concept_db)
