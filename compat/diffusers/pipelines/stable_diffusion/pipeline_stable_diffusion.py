import torch


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """diffusers 0.19 `rescale_noise_cfg` (torch formula; the inference path fuses it into the step kernel)."""
    std_text = noise_pred_text.std(dim=list(range(1, noise_pred_text.ndim)), keepdim=True)
    std_cfg = noise_cfg.std(dim=list(range(1, noise_cfg.ndim)), keepdim=True)
    noise_pred_rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * noise_pred_rescaled + (1 - guidance_rescale) * noise_cfg
