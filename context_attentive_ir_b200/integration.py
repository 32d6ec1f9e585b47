"""Drop-in hook: rebinds the network classes that the reference's model wrappers import by name
(neuroir/models/ranker.py:12-19, neuroir/models/multitask.py:12-14), so that the unchanged
main/ranker.py / main/multitask.py construct the B200 networks.  See INTEGRATION.md."""
from . import multitask, rankers


def install():
    """Call once before neuroir.models.Ranker / Multitask are instantiated.  Returns the list of rebound names."""
    done = []
    import neuroir.models.ranker as ref_ranker
    for name, cls in (('ARCI', rankers.ARCI), ('ARCII', rankers.ARCII), ('DSSM', rankers.DSSM), ('CDSSM', rankers.CDSSM), ('ESM', rankers.ESM), ('MatchTensor', rankers.MatchTensor), ('DRMM', rankers.DRMM),
                      ('DUET', rankers.DUET)):
        setattr(ref_ranker, name, cls)
        done.append('neuroir.models.ranker.' + name)
    try:
        import neuroir.models.multitask as ref_multitask
        for name in ('CARS', 'MNSRF', 'M_MATCH_TENSOR'):
            setattr(ref_multitask, name, getattr(multitask, name))
            done.append('neuroir.models.multitask.' + name)
    except ImportError:  # the multitask wrapper pulls in optional deps (prettytable, tqdm)
        pass
    return done
