% dump_goldens.m -- write real-reference goldens for the hot path (run in Octave/MATLAB with the reference
% magnusgrandin/ca-lanczos on the path; neither is available in the development image, so the committed fixtures in
% tests/golden/*.npz were generated from the oracle by tests/golden/make_golden.py instead).
%
%   octave --eval "addpath('/path/to/ca-lanczos'); run('tools/dump_goldens.m')"
%
% Writes tests/golden/ref_<name>.mat with T, the top Ritz values, the Newton shifts diag(Bk), sampled rows of the first
% basis block V and of Q, and the second-pass pattern -- the same quantities as the .npz fixtures.
function dump_goldens()
  dump_one('c1_poisson_s4_monomial', gallery('poisson', 100), 4, 60, 'monomial');
  n = 20000; dump_one('c2_diag_s8_newton', sparse(diag(linspace(1, 100, n))), 8, 64, 'newton');
end

function dump_one(name, A, s, iters, basis)
  n = size(A, 1); r = ones(n, 1);
  [T, Q] = ca_lanczos(A, r, s, iters, basis, 'local');
  ritz = sort(real(eig(T)), 'descend');
  q = r / sqrt(r' * r);
  if strcmpi(basis, 'newton')
    T0 = lanczos(A, q, 2 * s, 'full');
    shifts = leja(eig(T0), 'nonmodified');
    Bk = newton_basis_matrix(shifts, s, 1);
    V = matrix_powers_newton(A, q, s, diag(Bk), 1);
    shifts = diag(Bk);
  else
    V = [q, matrix_powers_monomial(A, q, s)];
    shifts = zeros(s, 1);
  end
  rows = unique(floor(linspace(0, n - 1, 257))) + 1;
  V_rows = V(rows, :); Q_rows = Q(rows, 1:2 * s + 1);
  save('-v7', fullfile('tests', 'golden', ['ref_' name '.mat']), 'T', 'ritz', 'shifts', 'rows', 'V_rows', 'Q_rows', 's', 'iters');
  printf('%s: ritz(1:3) = %.12f %.12f %.12f\n', name, ritz(1), ritz(2), ritz(3));
end
