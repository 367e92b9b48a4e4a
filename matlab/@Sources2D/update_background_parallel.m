function update_background_parallel(obj, use_parallel) %#ok<INUSD>
%% drop-in for ca_source_extraction/@Sources2D/update_background_parallel.m (ring model, bg_ssub = 1) on B200.
% Same signature and side effects (obj.W, obj.b0, obj.b0_new, obj.A_prev, obj.C_prev, log line); the arithmetic runs in
% libcnmfe_b200.so through cnmfe_b200_mex.  use_parallel is accepted and ignored: the library does its own intra-call
% concurrency (a parfor pool would create one CUDA context and one video copy per worker).
h = cnmfe_b200_handle(obj);                       % creates the context + uploads the blocks on first use
cnmfe_b200_mex('set_neurons', h, obj.A, obj.C);
o = obj.options;
cnmfe_b200_mex('set_options', h, cnmfe_b200_alg(o.spatial_algorithm), o.maxIter, o.deconv_flag, o.bg_acceleration, ...
    cnmfe_b200_deconv(o.deconv_options));
[r_shift, c_shift] = cnmfe_b200_mex('ring_offsets', h);
patch_pos = obj.P.mat_data.patch_pos;  block_pos = obj.P.mat_data.block_pos;
for m = 1:numel(patch_pos)
    cnmfe_b200_mex('set_ring', h, m-1, cnmfe_b200_W2slots(obj.W{m}, patch_pos{m}, block_pos{m}, r_shift, c_shift, obj.options), obj.b0{m});
end
cnmfe_b200_mex('update_background', h);
for m = 1:numel(patch_pos)
    tp = patch_pos{m};  dp = (diff(tp(1:2))+1) * (diff(tp(3:4))+1);
    [Ws, b0] = cnmfe_b200_mex('get_ring', h, m-1, numel(r_shift), dp);
    obj.W{m} = cnmfe_b200_slots2W(Ws, tp, block_pos{m}, r_shift, c_shift, obj.options);
    obj.b0{m} = b0;
end
obj.b0_new = obj.reconstruct_b0();
obj.A_prev = obj.A;
obj.C_prev = obj.C;
flog = fopen(obj.P.log_file, 'a');
fprintf(flog, '[%s]\b', get_minute());
fprintf(flog, 'Finished updating background using %s model.\n', obj.options.background_model);
fclose(flog);
end
