"""Replay of the reference's evaluation script on this package.

``test_mod_siren.py`` itself cannot run on the GPU box: /root/reference is not there, and its unrelated import-time
dependencies (polars, fastmri, scikit-image, matplotlib, seaborn) are not installed anywhere.  ``replay()`` therefore
restates the script's metric path LINE FOR LINE -- model construction (test_mod_siren.py:96-114), strict
``load_state_dict`` of a checkpoint file (:116-118), ``.to(device)`` (:120), the sample loop (:193-234) and the CSV
(:237-243) -- with the SAME import statements the script uses, resolved through ``mri_inr_b200.compat.install()``:

    from src.networks.modulated_siren import ModulatedSiren      # test_mod_siren.py:14
    from src.util.error import metrics_error                     # :15
    from src.util.tiling import image_to_patches                 # :16

Only the sampler is a stand-in (``SyntheticSampler``: same ``get_random_sample()`` / ``__len__`` contract as
``MRISampler``, src/data/mri_sampler.py:61-90, over synthetic slices instead of .npy files), and the plots
(:245-255, matplotlib/seaborn) are left out.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# configuration/test_modulated_siren.yaml:17-32 (model), :33-36 (testing), :15-16 (data)
BASELINE_CONFIG = dict(
    data=dict(dataset="synthetic", metric_samples=None, visual_samples=0, test_files=None, acceleration=6,
              center_fraction=0.05),
    model=dict(dim_in=2, dim_hidden=256, dim_out=1, latent_dim=256, num_layers=5, w0=1.0, w0_initial=30.0,
               use_bias=True, dropout=0.1, encoder_type="custom", encoder_path=None, outer_patch_size=32,
               inner_patch_size=16, siren_patch_size=24, activation="sine"),
    testing=dict(output_dir="./output", output_name="modulated_sired", model_path=None),
)


def make_config(**overrides) -> SimpleNamespace:
    """Nested namespace like ``load_configuration(path, testing=True)`` returns (src/configuration/configuration.py:164-185)."""
    cfg = {k: dict(v) for k, v in BASELINE_CONFIG.items()}
    for section, values in overrides.items():
        cfg[section].update(values)
    return SimpleNamespace(**{k: SimpleNamespace(**v) for k, v in cfg.items()})


class SyntheticSampler:
    """``MRISampler`` contract (mri_sampler.py:61-90): round-robin ``(fully_sampled [H,W], undersampled [H,W],
    slice_id)`` CPU tensors; ``len()`` = number of slices."""

    def __init__(self, fully_sampled: torch.Tensor, undersampled: torch.Tensor, names=None):
        assert fully_sampled.shape == undersampled.shape and fully_sampled.dim() == 3
        self.full, self.under = fully_sampled.cpu(), undersampled.cpu()
        self.names = list(names) if names is not None else [f"file_brain_AXFLAIR_synth_{i}" for i in range(len(self.full))]
        self.index_counter = 0

    def get_random_sample(self):
        idx = self.index_counter % len(self.full)
        self.index_counter += 1
        return self.full[idx], self.under[idx], self.names[idx]

    def __len__(self):
        return len(self.full)


def replay(config: SimpleNamespace, sampler) -> str:
    """test_mod_siren.py:80-243 (metric path).  Returns the path of ``metrics_error.csv``."""
    import mri_inr_b200.compat as compat

    compat.install(gpu_metrics=True)
    # --- the script's own import lines (test_mod_siren.py:14-16)
    from src.networks.modulated_siren import ModulatedSiren
    from src.util.error import metrics_error
    from src.util.tiling import image_to_patches

    output_dir = f"{config.testing.output_dir}/{config.testing.output_name}/test"          # :87
    os.makedirs(output_dir, exist_ok=True)

    # Set the device                                                                        # :89-93
    if torch.cuda.is_available():
        device = torch.device("cuda")
    else:
        device = torch.device("cpu")

    # Load the model                                                                        # :95-114
    mod_siren = ModulatedSiren(
        dim_in=config.model.dim_in,
        dim_hidden=config.model.dim_hidden,
        dim_out=config.model.dim_out,
        num_layers=config.model.num_layers,
        latent_dim=config.model.latent_dim,
        w0=config.model.w0,
        w0_initial=config.model.w0_initial,
        use_bias=config.model.use_bias,
        dropout=config.model.dropout,
        modulate=True,
        encoder_type=config.model.encoder_type,
        encoder_path=config.model.encoder_path,
        outer_patch_size=config.model.outer_patch_size,
        inner_patch_size=config.model.inner_patch_size,
        siren_patch_size=config.model.siren_patch_size,
        device=device,
        activation=config.model.activation,
    )

    mod_siren.load_state_dict(                                                              # :116-118
        torch.load(config.testing.model_path, map_location=device)
    )

    mod_siren.to(device)                                                                    # :120

    if not config.data.metric_samples:                                                      # :181-182
        config.data.metric_samples = len(sampler)

    if config.data.metric_samples > 0:                                                      # :184
        psnr_values = []
        ssim_values = []
        nrmse_values = []
        filenames = []

        with torch.no_grad():                                                               # :193-194
            mod_siren.eval()

            for i in range(config.data.metric_samples):                                     # :196
                # Load the image                                                            # :200-203
                fully_sampled_img, undersampled_img, filename = (
                    sampler.get_random_sample()
                )

                # unsqueeze image to add batch dimension                                    # :205-207
                fully_sampled_img = fully_sampled_img.unsqueeze(0).float().to(device)
                undersampled_img = undersampled_img.unsqueeze(0).float().to(device)

                fully_sampled_patch, _ = image_to_patches(                                  # :209-213
                    fully_sampled_img,
                    config.model.outer_patch_size,
                    config.model.inner_patch_size,
                )
                undersampled_patch, undersampled_information = image_to_patches(            # :214-218
                    undersampled_img,
                    config.model.outer_patch_size,
                    config.model.inner_patch_size,
                )

                psnr, ssim, nrmse = metrics_error(                                          # :220-229
                    mod_siren,
                    fully_sampled_patch,
                    undersampled_patch,
                    undersampled_information,
                    device,
                    config.model.outer_patch_size,
                    config.model.inner_patch_size,
                    config.model.siren_patch_size,
                )

                psnr_values.append(psnr)                                                    # :231-234
                ssim_values.append(ssim)
                nrmse_values.append(nrmse)
                filenames.append(filename)

            # Write them to a csv file                                                      # :236-243
            with open(os.path.join(output_dir, "metrics_error.csv"), "w") as f:
                f.write("FILENAME,PSNR,SSIM,NRMSE\n")
                for filename, psnr_value, ssim_value, nrmse_value in zip(
                    filenames, psnr_values, ssim_values, nrmse_values
                ):
                    f.write(f"{filename},{psnr_value},{ssim_value},{nrmse_value}\n")
    return os.path.join(output_dir, "metrics_error.csv")
