% calz_vec -- an n-by-cols real block that lives on the GPU (handle mode of the CA-Lanczos MEX gateways).
%
%   v = calz_vec(X)            upload a host matrix
%   V = matrix_powers_newton(A, v, s, lambda, 1)      % v a calz_vec  =>  V a calz_vec, nothing crosses PCIe
%   [QZ, RZ] = projectAndNormalize({Qprev}, V(:,2:s+1), true)   % calz_vec in, calz_vec out; RZ are host arrays
%   q = QZ(:,s);  Qprev = [q_handle ...]  -- column VIEWS share the storage (V(:,a:b), contiguous ranges only)
%   X = gather(v)              bring a block (or a view) back to the host
%
% The drivers (ca_lanczos.m, restarted_ca_lanczos.m) index Q as Q(:,a:b) and pass cell arrays of blocks: those
% expressions work unchanged on calz_vec objects; only the few lines that allocate Q = zeros(n,...) and assign blocks
% into it need the handle-aware variant shown in INTEGRATION.md.
classdef calz_vec < handle
    properties
        h = uint64(0);      % libcalz handle (calz_vec*); owned by the object that created it
        n = 0;
        col0 = 0;           % view: columns [col0, col0+cols) of the block
        cols = 0;
        owner = [];         % views keep their parent alive
    end
    methods
        function v = calz_vec(a, n, cols)
            if nargin == 3                       % (handle, n, cols): used by the gateways
                v.h = a; v.n = n; v.cols = cols;
            elseif nargin == 1                   % upload a host matrix
                v.n = size(a, 1); v.cols = size(a, 2);
                v.h = calz_vec_mex('create', v.n, v.cols);
                calz_vec_mex('upload', v.h, 0, full(double(a)));
            end
        end
        function w = subsref(v, S)
            if strcmp(S(1).type, '()') && numel(S(1).subs) == 2 && ischar(S(1).subs{1}) && strcmp(S(1).subs{1}, ':')
                idx = S(1).subs{2};
                if ischar(idx), idx = 1:v.cols; end
                if any(diff(idx) ~= 1), error('calanczos:badarg', 'calz_vec views are contiguous column ranges'); end
                w = calz_vec(v.h, v.n, numel(idx));
                w.col0 = v.col0 + idx(1) - 1;
                w.owner = v;
                if numel(S) > 1, w = subsref(w, S(2:end)); end
            else
                w = builtin('subsref', v, S);
            end
        end
        function v = subsasgn(v, S, rhs)         % Q(:,a:b) = Q_   (device-to-device, or an upload if rhs is a host matrix)
            if strcmp(S(1).type, '()') && numel(S) == 1 && numel(S(1).subs) == 2 && ischar(S(1).subs{1})
                idx = S(1).subs{2};
                if ischar(idx), idx = 1:v.cols; end
                if any(diff(idx) ~= 1), error('calanczos:badarg', 'calz_vec views are contiguous column ranges'); end
                if isa(rhs, 'calz_vec')
                    calz_vec_mex('copy', v.h, v.col0 + idx(1) - 1, rhs.h, rhs.col0, numel(idx));
                else
                    calz_vec_mex('upload', v.h, v.col0 + idx(1) - 1, full(double(rhs)));
                end
            else
                v = builtin('subsasgn', v, S, rhs);
            end
        end
        function varargout = size(v, d)
            sz = [v.n, v.cols];
            if nargin == 2, varargout{1} = sz(d);
            elseif nargout <= 1, varargout{1} = sz;
            else, varargout{1} = sz(1); varargout{2} = sz(2); end
        end
        function X = gather(v)
            X = calz_vec_mex('download', v.h, v.col0, v.cols);
        end
        function X = double(v), X = gather(v); end
        function delete(v)
            if isempty(v.owner) && v.h ~= 0, calz_vec_mex('free', v.h); end
        end
    end
end
