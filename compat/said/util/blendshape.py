from said_b200.util.blendshape import (  # noqa: F401
    load_blendshape_coeffs,
    save_blendshape_coeffs,
    save_blendshape_coeffs_batch,
    save_blendshape_coeffs_image,
)
