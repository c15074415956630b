from .models import NOF, Embedding, NOF_fine, NOF_coarse, NOF_plusfine, set_default_mlp_precision
