function h = cnmfe_b200_push(obj, what)
%% send the host state the next update reads to the library: ALWAYS the options (obj.options may have changed between calls)
% and the background state (obj.W / obj.b0, or obj.b / obj.f / obj.b0 -- they may have been reloaded from a saved run), plus
% the parts named in `what`: 'neurons' (obj.A, obj.C), 'prev' (obj.A_prev, obj.C_prev), 'sn' (obj.P.sn).
h = cnmfe_b200_handle(obj);
o = obj.options;
s = struct('thresh_outlier', o.thresh_outlier, 'spatial_algorithm', cnmfe_b200_alg(o.spatial_algorithm), 'maxIter', o.maxIter, 'deconv_flag', logical(o.deconv_flag), ...
    'bg_acceleration', logical(o.bg_acceleration), 'background_model', lower(o.background_model), 'nb', o.nb, 'bg_ssub', o.bg_ssub, ...
    'deconv_options', cnmfe_b200_deconv(o.deconv_options));
cnmfe_b200_mex('set_options', h, s);
md = obj.P.mat_data;
patch_pos = md.patch_pos;  block_pos = md.block_pos;
dims = md.dims;
if strcmpi(o.background_model, 'ring')
    for m = 1:numel(patch_pos)
        if isempty(obj.W) || isempty(obj.W{m}); continue; end
        if o.bg_ssub == 1
            [r_shift, c_shift] = cnmfe_b200_mex('ring_offsets', h);
            Ws = cnmfe_b200_W2slots(obj.W{m}, patch_pos{m}, block_pos{m}, r_shift, c_shift, dims(1), dims(2));
        else
            [d1s, d2s, ~, r_shift, c_shift] = cnmfe_b200_mex('ssub_dims', h, m-1);
            g = [1, d1s, 1, d2s];
            Ws = cnmfe_b200_W2slots(obj.W{m}, g, g, r_shift, c_shift, d1s, d2s);
        end
        cnmfe_b200_mex('set_ring', h, m-1, Ws, obj.b0{m});
    end
else
    for m = 1:numel(patch_pos)
        if isempty(obj.b) || isempty(obj.b{m}); continue; end
        b0 = [];
        if ~isempty(obj.b0) && ~isempty(obj.b0{m}); b0 = obj.b0{m}; end
        cnmfe_b200_mex('set_bf', h, m-1, full(obj.b{m}), full(obj.f{m}), b0);
    end
end
if any(strcmp(what, 'neurons')); cnmfe_b200_mex('set_neurons', h, obj.A, obj.C); end
if any(strcmp(what, 'prev'));    cnmfe_b200_mex('set_prev', h, obj.A_prev, obj.C_prev); end
if any(strcmp(what, 'sn'));      cnmfe_b200_mex('set_sn', h, obj.P.sn); end
end
