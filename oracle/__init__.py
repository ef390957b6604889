"""CPU oracle of the reference VAE-GAN step - test infrastructure only (see vaegan_oracle.py)."""
