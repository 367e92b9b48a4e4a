function update_background_parallel(obj, use_parallel) %#ok<INUSD>
%% drop-in for ca_source_extraction/@Sources2D/update_background_parallel.m on B200: ring model (bg_ssub >= 1), svd, nmf.
% Same signature and side effects (obj.W, obj.b0, obj.b, obj.f, obj.b0_new, obj.A_prev, obj.C_prev, log line, intermediate
% snapshot); the arithmetic runs in libcnmfe_b200.so through cnmfe_b200_mex.  use_parallel is accepted and ignored: the library
% does its own intra-call concurrency (a parfor pool would create one CUDA context and one video copy per worker).
what = {'neurons'};
if ~isnan(obj.options.thresh_outlier); what{end+1} = 'sn'; end   % the outlier clamp compares with thresh_outlier * obj.P.sn
h = cnmfe_b200_push(obj, what);                   % context + blocks on first use; options, W/b0 (or b/f), A, C
cnmfe_b200_mex('update_background', h);
md = obj.P.mat_data;
patch_pos = md.patch_pos;  block_pos = md.block_pos;  dims = md.dims;
T = diff(obj.frame_range) + 1;
if strcmpi(obj.options.background_model, 'ring')
    for m = 1:numel(patch_pos)
        tp = patch_pos{m};  dp = (diff(tp(1:2))+1) * (diff(tp(3:4))+1);
        if obj.options.bg_ssub == 1
            [r_shift, c_shift] = cnmfe_b200_mex('ring_offsets', h);
            [Ws, b0] = cnmfe_b200_mex('get_ring', h, m-1, numel(r_shift), dp, dp);
            obj.W{m} = cnmfe_b200_slots2W(Ws, tp, block_pos{m}, r_shift, c_shift, dims(1), dims(2));
        else
            [d1s, d2s, ~, r_shift, c_shift] = cnmfe_b200_mex('ssub_dims', h, m-1);
            [Ws, b0] = cnmfe_b200_mex('get_ring', h, m-1, numel(r_shift), d1s*d2s, dp);
            g = [1, d1s, 1, d2s];
            obj.W{m} = cnmfe_b200_slots2W(Ws, g, g, r_shift, c_shift, d1s, d2s);
        end
        obj.b0{m} = b0;
    end
else
    for m = 1:numel(patch_pos)
        tp = patch_pos{m};  dp = (diff(tp(1:2))+1) * (diff(tp(3:4))+1);
        [obj.b{m}, obj.f{m}, b0] = cnmfe_b200_mex('get_bf', h, m-1, dp, obj.options.nb, T);
        if strcmpi(obj.options.background_model, 'svd'); obj.b0{m} = b0; end   % fit_nmf_model returns no b0 (:231-236)
    end
end
obj.b0_new = obj.reconstruct_b0();
obj.A_prev = obj.A;
obj.C_prev = obj.C;
flog = fopen(obj.P.log_file, 'a');
fprintf(flog, '[%s]\b', get_minute());
fprintf(flog, 'Finished updating background using %s model.\n', obj.options.background_model);
cnmfe_b200_save_intermediate(obj, flog, 'bg', struct('b', {obj.b}, 'f', {obj.f}, 'b0', {obj.b0}, 'W', {obj.W}));
fclose(flog);
end
