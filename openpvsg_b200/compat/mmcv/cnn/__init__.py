def fuse_conv_bn(module):
    """mmcv.cnn.fuse_conv_bn: the B200 backbone already folds every eval-mode BatchNorm into its convolution when
    the weights are prepared (openpvsg_b200/mask2former.py::_fold), so there is nothing left to fuse."""
    return module
