class StableDiffusionSafetyChecker:
    """Placeholder: the reference only stores / monkey-patches this class
    (reference neuron_receivers/base_receiver.py:20-23)."""

    def forward(self, clip_input, images):
        return images, [False for _ in images]
