"""Inference CLI with the reference's command line (`script/inference.py` of SAiD): audio file -> blendshape CSV.

Same arguments, defaults and outputs as the reference script (argparse block `script/inference.py:20-118`,
flow `:139-214`); it imports the model and helpers through the same names (`said.model.diffusion`,
`said.util.*`, `dataset.dataset_voca`, `diffusers.DDIMScheduler`), which `compat/` maps onto this repository.

    python script/infer_cli.py --weights_path SAiD.pth --audio_path a.wav --output_path out.csv
"""
import argparse
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(_ROOT, "compat"), _ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
from dataset.dataset_voca import BlendVOCADataset  # noqa: E402
from diffusers import DDIMScheduler  # noqa: E402
from said.model.diffusion import SAID_UNet1D  # noqa: E402
from said.util.audio import fit_audio_unet, load_audio  # noqa: E402
from said.util.blendshape import load_blendshape_coeffs, save_blendshape_coeffs, save_blendshape_coeffs_image  # noqa: E402


# (flag, type, default) -- names, types and defaults of the reference CLI (script/inference.py:20-118)
FLAGS = [
    ("weights_path", str, "../BlendVOCA/SAiD.pth"),
    ("audio_path", str, "../BlendVOCA/audio/FaceTalk_170731_00024_TA/sentence01.wav"),
    ("output_path", str, "../out.csv"),
    ("output_image_path", str, "../out.png"),
    ("intermediate_dir", str, "../interm"),
    ("prediction_type", str, "epsilon"),
    ("save_image", bool, False),          # NB: as in the reference, any non-empty string is truthy
    ("save_intermediate", bool, False),
    ("num_steps", int, 1000),
    ("strength", float, 1.0),
    ("guidance_scale", float, 2.0),
    ("guidance_rescale", float, 0.0),
    ("eta", float, 0.0),
    ("fps", int, 60),
    ("divisor_unet", int, 1),
    ("unet_feature_dim", int, -1),
    ("device", str, "cuda:0"),
    ("init_sample_path", str, None),
    ("mask_path", str, None),
]


def main():
    parser = argparse.ArgumentParser(description="Audio file -> ARKit blendshape CSV with the B200-native SAiD engine")
    for name, typ, default in FLAGS:
        parser.add_argument("--" + name, type=typ, default=default)
    args = parser.parse_args()

    device = args.device
    init_samples = None
    if args.init_sample_path is not None:
        init_samples = load_blendshape_coeffs(args.init_sample_path).unsqueeze(0).to(device)
    mask = None
    if args.mask_path is not None:
        mask = load_blendshape_coeffs(args.mask_path).unsqueeze(0).to(device)

    said_model = SAID_UNet1D(noise_scheduler=DDIMScheduler, feature_dim=args.unet_feature_dim, prediction_type=args.prediction_type)
    said_model.load_state_dict(torch.load(args.weights_path, map_location=device))
    said_model.to(device)
    said_model.eval()

    waveform = load_audio(args.audio_path, said_model.sampling_rate)
    fit_output = fit_audio_unet(waveform, said_model.sampling_rate, args.fps, args.divisor_unet)
    waveform = fit_output.waveform
    window_len = fit_output.window_size
    waveform_processed = said_model.process_audio(waveform).to(device)

    with torch.no_grad():
        output = said_model.inference(
            waveform_processed=waveform_processed,
            init_samples=init_samples,
            mask=mask,
            num_inference_steps=args.num_steps,
            strength=args.strength,
            guidance_scale=args.guidance_scale,
            guidance_rescale=args.guidance_rescale,
            eta=args.eta,
            save_intermediate=args.save_intermediate,
            show_process=True,
        )

    result = output.result[0, :window_len].cpu().numpy()
    save_blendshape_coeffs(coeffs=result, classes=BlendVOCADataset.default_blendshape_classes, output_path=args.output_path)
    if args.save_image:
        save_blendshape_coeffs_image(result, args.output_image_path)
    if args.save_intermediate:
        os.makedirs(args.intermediate_dir, exist_ok=True)
        for t, interm in enumerate(reversed(output.intermediates)):
            interm_coeffs = interm[0, :window_len].cpu().numpy()
            timestep = t + 1
            save_blendshape_coeffs_image(interm_coeffs, os.path.join(args.intermediate_dir, f"{timestep}.png"))
            save_blendshape_coeffs(coeffs=interm_coeffs, classes=BlendVOCADataset.default_blendshape_classes,
                                   output_path=os.path.join(args.intermediate_dir, f"{timestep}.csv"))


if __name__ == "__main__":
    main()
