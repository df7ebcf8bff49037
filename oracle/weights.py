"""Seeded re-init recipe and synthetic images (re-exported from demonet_b200.seeded so that the oracle, the
golden generator and the product-side benchmark all draw the same tensors).  TEST INFRASTRUCTURE."""
from demonet_b200.seeded import DEFAULT_SEED, seeded_state_dict, seeded_vgg_state_dict, synthetic_images  # noqa: F401
