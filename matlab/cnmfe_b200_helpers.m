%% helper functions used by the drop-in methods (one file per function in a real installation).
%
% function h = cnmfe_b200_handle(obj)
%   if ~isfield(obj.P, 'b200_handle') || isempty(obj.P.b200_handle)
%       md = obj.P.mat_data;  dims = md.dims;
%       pp = int32(cell2mat(cellfun(@(x) x(:), md.patch_pos(:)', 'UniformOutput', false)));   % 4 x npatch, MATLAB linear order
%       bp = int32(cell2mat(cellfun(@(x) x(:), md.block_pos(:)', 'UniformOutput', false)));
%       T = diff(obj.frame_range) + 1;
%       h = cnmfe_b200_mex('create', dims(1), dims(2), T, pp, bp, obj.options.ring_radius, ...
%                          max(0, double(obj.options.num_neighbors)), 0);
%       for m = 1:numel(md.patch_pos)
%           Yb = get_patch_data(md, md.patch_pos{m}, obj.frame_range, true);    % native dtype, with overlap
%           cnmfe_b200_mex('upload_block', h, m-1, Yb);
%       end
%       obj.P.b200_handle = h;
%   end
%   h = obj.P.b200_handle;
%
% function Ws = cnmfe_b200_W2slots(W, tp, tb, r_shift, c_shift, options)      % sparse (d_p x d_blk) -> nnb x d_p
%   nr = diff(tp(1:2))+1; nc = diff(tp(3:4))+1; nrb = diff(tb(1:2))+1;
%   [cc, rr] = meshgrid(tp(3):tp(4), tp(1):tp(2));  rr = rr(:);  cc = cc(:);
%   Ws = zeros(numel(r_shift), nr*nc);
%   for s = 1:numel(r_shift)
%       r2 = rr + double(r_shift(s));  c2 = cc + double(c_shift(s));
%       ok = r2>=1 & r2<=options.d1 & c2>=1 & c2<=options.d2;
%       jj = (c2-tb(3))*nrb + (r2-tb(1)+1);
%       idx = sub2ind(size(W), find(ok), jj(ok));
%       Ws(s, ok) = full(W(idx));
%   end
%
% function W = cnmfe_b200_slots2W(Ws, tp, tb, r_shift, c_shift, options)      % inverse of the above
%   (same index computation, W = sparse(ii, jj, vals, nr*nc, nrb*ncb))
%
% function a = cnmfe_b200_alg(name)      % 'hals'->0, 'hals_thresh'->1, 'nnls'->2, 'lars'->3
% function s = cnmfe_b200_deconv(opts)   % run deconvolveCa's own parseinputs on the cell/struct -> struct
