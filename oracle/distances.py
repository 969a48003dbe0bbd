"""Oracle: distances.  Follows /root/reference/src/stats/distances.jl:29-31, 51-53."""
import numpy as np


def linear_distance(data, synth_data):
    """mean(abs2.(data .- synth_data))  (distances.jl:29-31)."""
    a = np.asarray(data, dtype=np.float64)
    b = np.asarray(synth_data, dtype=np.float64)
    return float(np.mean((a - b) ** 2))


def logarithmic_distance(data, synth_data):
    """mean(abs2.(log.(data) .- log.(synth_data)))  (distances.jl:51-53).  Julia's
    log throws DomainError on negative reals; on the batched path a non-positive
    statistic yields NaN (never accepted, abc.jl:185)."""
    a = np.asarray(data, dtype=np.float64)
    b = np.asarray(synth_data, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.mean((np.log(a) - np.log(b)) ** 2))
