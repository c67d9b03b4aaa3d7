"""CPU oracle for segmantic's volumetric prediction path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``segmantic_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and only as the checker / the timed CPU baseline.

PARITY UNPINNED.  The reference (``/root/reference``, dyollb/segmantic 0.4.0) composes its hot path
from third-party libraries that are neither vendored in the reference tree nor installable in this
image (no network): MONAI (unpinned in ``pyproject.toml:29``; effective >= 1.2, ~1.3.x),
pytorch-lightning (``pyproject.toml:31``) and SimpleITK/ITK (``pyproject.toml:33``).  The reference's
own tests hold no golden vector for this path (``tests/seg/test_unet.py:15-20`` checks hparams only;
``tests/image/test_image.py:33-52`` checks sizes/spacings only).  This package therefore *restates*
the published algorithms of those libraries, anchored on the reference's call sites:

* ``unet.py``            MONAI ``UNet`` as configured at ``src/segmantic/seg/monai_unet.py:114-124``
* ``sliding_window.py``  MONAI ``SlidingWindowInferer`` as called at ``seg/monai_unet.py:637-639,665``
* ``spacing.py``         MONAI ``Spacingd`` / ``Invertd`` / ``Orientationd`` / ``NormalizeIntensityd`` /
                         ``CropForegroundd`` as configured at ``seg/monai_unet.py:151-176,612-625``
* ``itk_resample.py``    ITK ``ResampleImageFilter`` as configured at ``image/processing.py:49-120``
* ``predict.py``         the composition of the above = ``predict()`` at ``seg/monai_unet.py:551-670``
* ``bf16_emulation.py``  the same UNet with the device path's bf16 rounding points (checks the tensor-core path)
* ``evaluation.py``      ``confusion_matrix`` (``seg/evaluation.py:96-125``) and MONAI ``DiceMetric`` /
                         ``ConfusionMatrixMetric`` as used at ``seg/monai_unet.py:640-725``
* ``ensemble.py``        MONAI ``MeanEnsemble`` / ``VoteEnsemble`` and the reference's own ``SelectBestEnsemble``
                         (``seg/transforms.py:15-61``) as used at ``seg/monai_unet.py:917-1003``

The conv / grid_sample / argmax primitives underneath are the *real* ``torch`` CPU kernels (torch is
installed), so layer-level arithmetic is executable-reference, only the composition is restated.
"""
