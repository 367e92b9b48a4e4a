function update_temporal_parallel(obj, use_parallel, use_c_hat) %#ok<INUSL>
%% drop-in for ca_source_extraction/@Sources2D/update_temporal_parallel.m on B200 (use_c_hat = false is not built).
if exist('use_c_hat', 'var') && ~isempty(use_c_hat) && ~use_c_hat
    error('cnmfe:b200', 'use_c_hat=false (fast_temporal) is not available in the B200 path');
end
h = cnmfe_b200_handle(obj);
cnmfe_b200_mex('set_neurons', h, obj.A, obj.C);
cnmfe_b200_mex('set_prev', h, obj.A_prev, obj.C_prev);
[K, T] = size(obj.C);
[C, C_raw, S, kp, nsn] = cnmfe_b200_mex('update_temporal', h, K, T);
obj.C = C;  obj.C_raw = C_raw;  obj.S = sparse(S);
p = 1 + strcmpi(obj.options.deconv_options.type, 'ar2');
obj.P.kernel_pars = kp(1:p, :)';
obj.P.neuron_sn = nsn;
if strcmpi(obj.options.background_model, 'ring')
    obj.b0_new = cell2mat(obj.P.Ymean) - obj.reshape(obj.A*mean(obj.C,2), 2);
end
flog = fopen(obj.P.log_file, 'a');
fprintf(flog, '[%s]\b', get_minute());
fprintf(flog, 'Finished updating temporal components.\n');
fclose(flog);
end
