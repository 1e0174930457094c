"""``Lens`` — orchestrator of the concept-database path; drop-in for ``semanticlens.lens`` (reference lens.py:27-480).

Same functions, signatures, cache-file grammar and control flow as the reference. What changes underneath:
``cv._compute_concept_db(fm)`` is the B200 sweep + embed + gather, and ``_probe`` / ``eval_*`` call the libslb200 score
kernels (``semanticlens_b200.scores``).
"""

from __future__ import annotations

import logging

import torch
from safetensors.torch import load_file, save_file
from tqdm.auto import tqdm

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
    """Reference lens.py:166-203, including its template layout: the templated list is built template-major
    (``for t in templates for q in query``) but regrouped as ``(q t)`` — kept as is."""
    if templates:
        query_templated = [t.format(q) for t in templates for q in query]
        empty_templates = [t.format("") for t in templates]
        batch_size = batch_size or len(query_templated)
        chunks = []
        for b0 in tqdm(range(0, len(query_templated), batch_size), desc="text embedding ...", leave=False,
                       disable=batch_size == len(query_templated)):
            batch = query_templated[b0 : b0 + batch_size]
            chunks.append(fm.encode_text(fm.tokenize(batch).to(fm.device)).cpu())
        templated = torch.cat(chunks, dim=0)
        empty = fm.encode_text(fm.tokenize(empty_templates).to(fm.device)).cpu()
        q, t = len(query), len(templates)
        return (templated.reshape(q, t, -1) - empty.reshape(1, t, -1)).mean(1)
    return fm.encode_text(fm.tokenize(query).to(fm.device)).cpu()


@torch.no_grad()
def _probe(query: torch.Tensor, aggregated_concept_db):
    """K6 behind the reference's tensor-or-dict dispatch (lens.py:207-214)."""
    if isinstance(aggregated_concept_db, torch.Tensor):
        return similarity_score(query.to(aggregated_concept_db.device), aggregated_concept_db)
    return {key: similarity_score(query.to(value.device), value) for key, value in aggregated_concept_db.items()}


class Lens:
    """Holds the foundation model; computes / caches concept DBs and evaluates them (reference lens.py:217-480)."""

    def __init__(self, fm, device=None):
        self.fm = fm
        self.device = device or self.fm.device
        self.fm.to(self.device)
        if not hasattr(self.fm, "name"):
            self.fm.name = get_fallback_name(self.fm)
            logger.debug(f"Assigned fallback name to foundation model: {self.fm.name}")

    def compute_concept_db(self, cv: AbstractComponentVisualizer, **kwargs) -> dict[str, torch.Tensor]:
        """Concept DB of ``cv`` under ``self.fm``, cached as safetensors next to the act-max cache when ``cv.caching``
        (file grammar of reference lens.py:308-316)."""
        if cv.caching:
            fdir = cv.storage_dir / "concept_database" / self.fm.name
            fdir.mkdir(parents=True, exist_ok=True)
            fname = (
                "concept_db-"
                + "-".join([v for k, v in cv.metadata.items() if k not in ["dataset", "model"]])
                + ".safetensors"
            )
            fpath = fdir / fname
            if fpath.exists():
                logger.debug("Loading concept DB from cache")
                return load_file(filename=fpath)
            logger.debug("Computing concept DB and saving to cache")
            concept_db = cv._compute_concept_db(self.fm, **kwargs)
            save_file(tensors={k: v.cpu().contiguous() for k, v in concept_db.items()}, filename=fpath)
            logger.debug(f"Saved concept DB to cache {fpath}")
            return concept_db
        logger.debug("Caching is not enabled. Computing Concept DB")
        return cv._compute_concept_db(self.fm, **kwargs)

    def text_probing(self, query, aggregated_concept_db, templates=None, batch_size=None):
        return text_probing(self.fm, query, aggregated_concept_db, templates, batch_size)

    def image_probing(self, query, aggregated_concept_db):
        return image_probing(self.fm, query, aggregated_concept_db)

    def eval_clarity(self, concept_db):
        if isinstance(concept_db, torch.Tensor):
            return clarity_score(concept_db)
        return {key: clarity_score(value) for key, value in concept_db.items()}

    def eval_redundancy(self, aggregated_concept_db):
        if isinstance(aggregated_concept_db, torch.Tensor):
            return redundancy_score(aggregated_concept_db)
        return {key: redundancy_score(value) for key, value in aggregated_concept_db.items()}

    def eval_polysemanticity(self, concept_db):
        if isinstance(concept_db, torch.Tensor):
            return polysemanticity_score(concept_db)
        return {key: polysemanticity_score(value) for key, value in concept_db.items()}
