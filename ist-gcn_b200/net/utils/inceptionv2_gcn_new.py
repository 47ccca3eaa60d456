"""``net.utils.inceptionv2_gcn_new`` of the reference is ``inceptionv2_gcn`` with the einsums inlined
(net/utils/inceptionv2_gcn_new.py:46-57) and without the unused ``t_*`` constructor arguments."""
from net.utils.inceptionv2_gcn import Inception2 as _Inception2


class Inception2(_Inception2):

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__(in_channels, out_channels, kernel_size)
