function h = cnmfe_b200_handle(obj)
%% context of the B200 library for this Sources2D object: created (and the video blocks uploaded) on first use.
% The handle (uint64) lives in obj.P.b200_handle -- Sources2D is a handle class, so it persists across calls.
% Geometry = what distribute_data.m:163-173 stored in mat_data (patch_pos / block_pos, 1-based inclusive [r0 r1 c0 c1]);
% blocks are read with the reference's own get_patch_data(mat_data, patch_pos, frame_range, true) (with overlap).
if isfield(obj.P, 'b200_handle') && ~isempty(obj.P.b200_handle)
    h = obj.P.b200_handle;
    return;
end
md = obj.P.mat_data;
dims = md.dims;
patch_pos = md.patch_pos;  block_pos = md.block_pos;
np = numel(patch_pos);
pp = zeros(4, np, 'int32');  bp = zeros(4, np, 'int32');
for m = 1:np                                     % MATLAB linear order over (nr_patch, nc_patch)
    pp(:, m) = int32(reshape(patch_pos{m}, 4, 1));
    bp(:, m) = int32(reshape(block_pos{m}, 4, 1));
end
T = diff(obj.frame_range) + 1;
nn = obj.options.num_neighbors;
if isempty(nn) || isnan(nn); nn = 0; end
dev = 0;
if isfield(obj.P, 'b200_device') && ~isempty(obj.P.b200_device); dev = obj.P.b200_device; end
h = cnmfe_b200_mex('create', dims(1), dims(2), T, pp, bp, ceil(obj.options.ring_radius), double(nn), dev);
for m = 1:np
    Yb = get_patch_data(md, patch_pos{m}, obj.frame_range, true);     % native dtype (uint8 / uint16), nr_block x nc_block x T
    cnmfe_b200_mex('upload_block', h, m-1, Yb);
end
obj.P.b200_handle = h;
end
