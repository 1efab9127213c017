// only the Misner angle is identified, with the period the user chooses (0: not identified)
(T, mphi, y, z) => [0, $cfg.phi0, 0, 0]
