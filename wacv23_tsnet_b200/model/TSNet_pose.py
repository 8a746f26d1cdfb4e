"""Drop-in `TSNet` (pose variant) -- surface of the reference model/TSNet_pose.py:206-417.

Identical generator; the output is composited with a fixed foreground mask (columns 64:192) over the
mean colour (model/TSNet_pose.py:276-280, 416-417), fused into the output-head kernel.
"""
import numpy as np
import torch

from . import TSNet as _face


class TSNet(_face.TSNet):
    def __init__(self, lr=0.0002, beta1=0.5, n_blocks=0,
                 n_source=3,
                 lambda_FML=10.0, lambda_VGG=10.0, lambda_CON=10.0, lambda_GRAD=10.0,
                 is_train=True, getIntermFeat=True, label_nc=5,
                 debug=False, lambda_dec=1.0,
                 addcoords=True,
                 ngf=64, n_downsampling=4,
                 use_mask=True,
                 mean=np.array((101.84807705937696, 112.10832843463207, 111.65973036298041), dtype=np.float32),
                 math_mode="fp16x3", cuda_graph=False, winograd=True):
        super().__init__(lr=lr, beta1=beta1, n_blocks=n_blocks, n_source=n_source, lambda_FML=lambda_FML,
                         lambda_VGG=lambda_VGG, lambda_CON=lambda_CON, lambda_GRAD=lambda_GRAD, is_train=is_train,
                         getIntermFeat=getIntermFeat, label_nc=label_nc, debug=debug, lambda_dec=lambda_dec,
                         addcoords=addcoords, ngf=ngf, n_downsampling=n_downsampling, return_flow=False,
                         math_mode=math_mode, cuda_graph=cuda_graph, winograd=winograd,
                         img_mean=np.asarray(mean, dtype=np.float32))
        self.model_names = ['G', 'D', 'DF']
        self.use_mask = use_mask
        if self.use_mask:
            # same arithmetic as the reference's fp32 (-mean) / 255.0, evaluated on the host: ATen's CUDA kernel for
            # tensor / python-scalar multiplies by the reciprocal (1 ulp off a true division), and the parity gate
            # is the reference's CPU forward.
            self.mask_img = (torch.from_numpy(-np.asarray(mean, dtype=np.float32)).view(1, 3, 1, 1)
                             .repeat(1, 1, 256, 256) / 255.0).cuda()
            fore_mask = torch.zeros((256, 256), dtype=torch.float32)
            fore_mask[:, 64:192] = 1
            self.fore_mask = fore_mask.view(1, 1, 256, 256).cuda()
            self._pose_fill = tuple(float(v) for v in self.mask_img[0, :, 0, 0].cpu())

    def crop_face(self, image, real_lbl):
        raise NotImplementedError(_face._TRAIN_MSG)
