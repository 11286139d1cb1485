"""``pix2pose_model.ae_model`` inference builders -> B200 drop-in (see pix2pose_b200/ae_model.py).
Training-only symbols of the reference (transformer_loss, DCGAN_discriminator) are out of scope."""
from pix2pose_b200.ae_model import aemodel_unet_prob, aemodel_unet_resnet50, GeneratorModel  # noqa: F401
