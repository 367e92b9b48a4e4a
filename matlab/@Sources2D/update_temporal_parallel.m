function update_temporal_parallel(obj, use_parallel, use_c_hat) %#ok<INUSL>
%% drop-in for ca_source_extraction/@Sources2D/update_temporal_parallel.m on B200.
% use_c_hat = false selects fast_temporal (:314-337); default true (HALS sweeps + deconvolution, :60-62).
if ~exist('use_c_hat', 'var') || isempty(use_c_hat); use_c_hat = true; end
h = cnmfe_b200_push(obj, {'neurons', 'prev'});
cnmfe_b200_mex('set_use_c_hat', h, logical(use_c_hat));
[K, T] = size(obj.C);
[C, C_raw, S, kp, nsn] = cnmfe_b200_mex('update_temporal', h, K, T);
obj.C = C;  obj.C_raw = C_raw;  obj.S = sparse(S);
if obj.options.deconv_flag
    dopt = cnmfe_b200_deconv(obj.options.deconv_options);
    p = 1;
    if isfield(dopt, 'type') && strcmpi(dopt.type, 'ar2'); p = 2; end
    obj.P.kernel_pars = kp(1:p, :)';               % deconvTemporal.m:100-105
    obj.P.neuron_sn = nsn;
end
if strcmpi(obj.options.background_model, 'ring')
    obj.b0_new = cell2mat(obj.P.Ymean) - obj.reshape(obj.A*mean(obj.C,2), 2);
end
flog = fopen(obj.P.log_file, 'a');
fprintf(flog, '[%s]\b', get_minute());
fprintf(flog, 'Finished updating temporal components.\n');
temporal = struct('C_raw', obj.C_raw, 'ids', obj.ids, 'C', obj.C, 'S', obj.S, 'b0_new', obj.b0_new);
temporal.P.kernel_pars = obj.P.kernel_pars;
cnmfe_b200_save_intermediate(obj, flog, 'temporal', temporal);
fclose(flog);
end
