"""GA3C actor -> predictor loop on the GPU, with the reference's Config / NetworkVP / Server surface."""
