class LearnedPerceptualImagePatchSimilarity:
    def __init__(self, *a, **k):
        pass

    def cuda(self):
        return self

    def __call__(self, *a, **k):
        raise NotImplementedError("LPIPS is not available offline")
