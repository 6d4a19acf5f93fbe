"""B200-native (sm_100a) perceptual-loss training step of edogariu/artist-style-transfer.

Drop-in mirrors of the reference modules: `cnn.StyleTransfer`, `train_cnn.VGG16`, `train_cnn.gram`.
All arithmetic is in libast_b200.so (C ABI: include/ast.h); importing this package never imports `oracle`.
"""
from . import _lib
from .cnn import (ConvLayer, DeconvLayer, ResidualLayer, StyleTransfer, TransformerNet, get_default_precision,
                  set_default_precision)
from .train_cnn import (VGG16, PerceptualTrainer, StyleGramBank, gram, mse_loss, neg_mean, perceptual_losses,
                        perceptual_step, style_grams_single, style_grams_smartaverage)

__all__ = ["StyleTransfer", "TransformerNet", "ConvLayer", "ResidualLayer", "DeconvLayer", "VGG16", "gram",
           "mse_loss", "perceptual_step", "perceptual_losses", "PerceptualTrainer", "StyleGramBank", "style_grams_single", "style_grams_smartaverage",
           "set_default_precision", "get_default_precision", "neg_mean"]
