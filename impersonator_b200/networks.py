"""``networks/networks.py:10-42``: the factory ``models/imitator.py`` builds its networks through
(``from networks.networks import NetworksFactory, HumanModelRecovery``)."""
from .generator import ImpersonatorGenerator, NetworkBase                   # noqa: F401
from .hmr import HumanModelRecovery                                          # noqa: F401


class NetworksFactory(object):
    def __init__(self):
        pass

    @staticmethod
    def get_by_name(network_name, *args, **kwargs):
        if network_name == 'impersonator':
            network = ImpersonatorGenerator(*args, **kwargs)
        elif network_name == 'deepfillv2':
            from .inpaintor import InpaintSANet
            network = InpaintSANet(*args, **kwargs)
        elif network_name in ('concat', 'discriminator_patch_gan', 'global_local'):
            raise ValueError("Network %s belongs to the baselines / training code and is not part of the B200 inference "
                             "hot path (SURVEY.md section 8)" % network_name)
        else:
            raise ValueError("Network %s not recognized." % network_name)
        print("Network %s was created" % network_name)
        return network
