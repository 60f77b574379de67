% dump_goldens.m -- write REAL-REFERENCE goldens for the hot path and its callers.  Run in Octave/MATLAB with the reference
% magnusgrandin/ca-lanczos on the path (neither is available in the development image, so the committed fixtures in
% tests/golden/*.npz were generated from the oracle by tests/golden/make_golden.py and parity is "unpinned" until this script
% has been run once):
%
%   octave --eval "addpath('/path/to/ca-lanczos'); run('tools/dump_goldens.m')"
%   python -m pytest tests/test_reference_goldens.py -q        # picks up tests/golden/ref_*.mat and pins the oracle
%
% Every input is a closed-form expression (no random numbers), restated in tests/test_reference_goldens.py.
function dump_goldens()
  outdir = fullfile('tests', 'golden');
  % ---- drivers: ca_lanczos 'local' (configs C1 and C2 of BASELINE.json)
  dump_driver(outdir, 'c1_poisson_s4_monomial', gallery('poisson', 100), 4, 60, 'monomial');
  n = 20000; dump_driver(outdir, 'c2_diag_s8_newton', sparse(diag(linspace(1, 100, n))), 8, 64, 'newton');
  % ---- kernels on closed-form blocks: tsqr, cholqr, normalize, project, projectAndNormalize (both branches of the 50 % test)
  n = 3000; c = 6; m = 7;
  [I, J] = ndgrid(1:n, 1:c);   X = cos(0.37 * I .* J + J) + 0.01 * I / n;
  [I, J] = ndgrid(1:n, 1:m);   B = sin(0.11 * I .* J + 2 * J) + 0.5 * cos(0.05 * I);
  [Qb, Rb] = tsqr(B);
  [Qt, Rt] = tsqr(X);
  [Qc, Rc] = cholqr(X);
  [Qn, Rn, rk] = normalize(X);
  [Yp, Rp] = project({Qb}, X);
  far = X;  near = Qb * ones(m, c) + 1e-2 * X;
  [QZf, RZf] = projectAndNormalize({Qb}, far, true);
  [QZn, RZn] = projectAndNormalize({Qb}, near, true);
  rows = unique(floor(linspace(0, n - 1, 129))) + 1;
  Qb_rows = Qb(rows, :); Qt_rows = Qt(rows, :); Qc_rows = Qc(rows, :); Yp_rows = Yp(rows, :);
  QZf_rows = QZf(rows, :); QZn_rows = QZn(rows, :);
  Rp1 = Rp{1}; RZf1 = RZf{1}; RZf2 = RZf{2}; RZn1 = RZn{1}; RZn2 = RZn{2};
  save('-v7', fullfile(outdir, 'ref_kernels.mat'), 'n', 'c', 'm', 'rows', 'Rb', 'Rt', 'Rc', 'Rn', 'rk', 'Rp1', 'RZf1', 'RZf2', 'RZn1', ...
       'RZn2', 'Qb_rows', 'Qt_rows', 'Qc_rows', 'Yp_rows', 'QZf_rows', 'QZn_rows');
  % ---- matrix powers kernel with a complex-conjugate shift pair (modified Newton basis)
  A = gallery('poisson', 30); q = ones(900, 1) / 30;
  lambda = [7.5; 1 + 2i; 1 - 2i; 4];
  Vc = matrix_powers_newton(A, q, 4, lambda, 1);
  save('-v7', fullfile(outdir, 'ref_mpk_complex_pair.mat'), 'Vc');
  % ---- restarted driver on the reference's own test (test_restart_diagonal_matrices.m:8-36, N scaled to 2000)
  N = 2000; A = sparse(diag(linspace(1, 1e2, N))); r = ones(N, 1);
  modes = {'local', 'full', 'periodic', 'selective'};
  for i = 1:numel(modes)
    [E, V, nres, rnorms, ortherr] = restarted_ca_lanczos(A, r, 40, 4, 4, 'newton', modes{i}, 1e-8);
    save('-v7', fullfile(outdir, ['ref_restart_' modes{i} '.mat']), 'E', 'nres', 'rnorms', 'ortherr');
  end
  % ---- periodic ca_lanczos on test_convergence_diagonal_matrices.m:9-22
  N = 500; A = sparse(diag(linspace(1, 100, N))); r = ones(N, 1);
  [T, Q] = ca_lanczos(A, r, 8, 480, 'newton', 'periodic');
  ritz = sort(real(eig(T)), 'descend'); ritz = ritz(1:20);
  save('-v7', fullfile(outdir, 'ref_periodic_diag500.mat'), 'ritz');
end

function dump_driver(outdir, name, A, s, iters, basis)
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
  save('-v7', fullfile(outdir, ['ref_' name '.mat']), 'T', 'ritz', 'shifts', 'rows', 'V_rows', 'Q_rows', 's', 'iters');
  printf('%s: ritz(1:3) = %.12f %.12f %.12f\n', name, ritz(1), ritz(2), ritz(3));
end
