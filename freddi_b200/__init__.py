"""freddi_b200 — B200 (sm_100a) ensemble engine for Freddi's disc-evolution hot path.

``Freddi`` / ``FreddiNeutronStar`` keep the reference's Python surface; ``Ensemble`` evolves many
independent models at once.  Importing the compute classes loads ``libfreddi_b200.so``; there is no CPU
fallback (``freddi_b200.args`` and ``freddi_b200.build`` are importable without it).
"""
__all__ = ('Freddi', 'FreddiNeutronStar', 'EvolutionResult', 'Ensemble', 'make_args', 'sweep', 'sweep_plan', 'pinned_empty')


def __getattr__(name):
    if name in ('Freddi', 'FreddiNeutronStar', 'EvolutionResult'):
        from . import freddi
        return getattr(freddi, name)
    if name in ('Ensemble', 'column_names', 'shard_indices', 'sweep', 'sweep_plan', 'pinned_empty', 'SweepResult'):
        from . import ensemble
        return getattr(ensemble, name)
    if name == 'make_args':
        from .args import make_args
        return make_args
    raise AttributeError(name)
