def Q_discrete_white_noise(*a, **k):
    raise NotImplementedError
