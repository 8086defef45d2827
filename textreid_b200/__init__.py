"""textreid_b200 -- B200-native hot path of BrandonHanx/TextReID: the MoCo cross-modal loss step and
the text-to-image retrieval evaluation, behind the reference's Python call shapes.

Importing the package does not need a GPU; calling anything does (there is no CPU fallback).
"""
from . import _lib
from .evaluation import (RetrievalResult, build_relevance, compute_on_dataset, compute_on_dataset_tensors, eval_precision,
                         evaluate_embeddings, evaluation, gather_embeddings, inference, l2_normalize_rows, rank,
                         rank_artifacts, retrieve)
from .sharded import ShardPlan, clear_plan_cache, retrieve_sharded, retrieve_sharded_local
from .losses import MomentumUpdater, dequeue_and_enqueue, ema_update_flat, moco_loss_dict
from .moco_head import FusedMoCoHead, LossComputation, build_moco_head

__all__ = [
    "FusedMoCoHead", "LossComputation", "build_moco_head", "moco_loss_dict", "dequeue_and_enqueue",
    "MomentumUpdater", "ema_update_flat", "rank", "rank_artifacts", "retrieve", "evaluation", "inference",
    "build_relevance", "l2_normalize_rows", "RetrievalResult", "compute_on_dataset", "compute_on_dataset_tensors",
    "gather_embeddings", "evaluate_embeddings", "eval_precision", "retrieve_sharded", "retrieve_sharded_local", "ShardPlan",
    "clear_plan_cache",
]
__version__ = "0.1.0"
