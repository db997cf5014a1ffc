from scipy.special import erf  # noqa: F401
