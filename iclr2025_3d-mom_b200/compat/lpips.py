"""stub: lpips_loss is never called on the 4DGS path (utils/loss_utils.py:16)."""
